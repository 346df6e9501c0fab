// solb_handles.h — the opaque handles of include/solb.h and the error plumbing, shared by solb_api.cu and comm.cu.
#pragma once
#include "../../include/solb.h"

#include <algorithm>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "solb_internal.h"
#include "trace.h"

using namespace solb;

extern thread_local std::string g_last_error;

struct solb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    std::string err;
    cudaMemPool_t build_pool = nullptr;     // private stream-ordered pool of the build scratch (release threshold raised)
    unsigned long long *d_stats = nullptr;  // 8 slots
    uint64_t launches = 0;
    bool timing = false;
    bool preloaded = false;  // solb_ctx_preload ran
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_build_ms = 0.0f, last_trace_ms = 0.0f;
    uint32_t *d_blue = nullptr;
    uint32_t blue_w = 0, blue_h = 0;
    uint32_t *pinned_count = nullptr;
    WavefrontState ws = {};
    // warp-local wavefront schedule: slot-indexed path state sized by the persistent grid, and the per-pixel frame sums the
    // kernel hands to the resolve; two of each plus two side streams, so that the kernel of frame f + 1 can start while frame
    // f drains (the resolves into the targets stay in order on the ctx stream)
    WarpfrontState wl[WL_MAX_FRAMES] = {};
    float4 *frame_sum[WL_MAX_FRAMES] = {};
    size_t frame_sum_pixels[WL_MAX_FRAMES] = {};
    cudaStream_t frame_stream[WL_MAX_FRAMES] = {};
    cudaEvent_t ev_trace_done[WL_MAX_FRAMES] = {}, ev_resolve_done[WL_MAX_FRAMES] = {}, ev_side_barrier = nullptr, ev_serial = nullptr;
    cudaEvent_t ev_k0[WL_MAX_FRAMES] = {}, ev_k1[WL_MAX_FRAMES] = {};  // timing mode: around the kernel on its side stream
    bool have_resolve_done[WL_MAX_FRAMES] = {}, have_side_barrier = false;
    uint32_t frame_slot = 0;
    // greatest-priority stream for the short kernels that follow a frame (resolve, band pack / all-gather / scatter): the
    // persistent trace kernels of the frames in flight fill every SM, and a kernel of their priority waits for a whole drain
    cudaStream_t hi_stream = nullptr;
    cudaEvent_t ev_hi_in = nullptr, ev_hi_out = nullptr;
    int use_hi_stream = 0;  // SOLB_HI_STREAM=1: measured neutral with the peer-store exchange, slower with the NCCL one (profiles/r02_tile_split.txt)
    // queues + counters + streams of the extra frame parts (overlap mode); index 0 unused (= ws / stream)
    uint32_t *part_queue[WF_MAX_PARTS][2] = {};
    uint32_t *part_counters[WF_MAX_PARTS] = {};
    cudaStream_t part_stream[WF_MAX_PARTS] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[WF_MAX_PARTS] = {}, ev_poll[WF_MAX_PARTS] = {};
    cudaStream_t shade_stream[WF_MAX_PARTS] = {};
    cudaEvent_t ev_ts[WF_MAX_PARTS] = {}, ev_st[WF_MAX_PARTS] = {};
    uint2 *part_spill[WF_MAX_PARTS] = {};  // ray-pool kernel stack spill, one per frame part
    std::vector<cudaEvent_t> ev_pool;  // timing mode: pairs around every dominant-kernel launch
    float trace_kernel_ms_total = 0.0f;
    uint32_t trace_kernel_launches = 0;
    TraceTuning tune;
    uint32_t auto_wide_schedule = SOLB_SCHEDULE_WARPFRONT;  // what SOLB_SCHEDULE_AUTO picks above 8 wide nodes (SOLB_AUTO_SCHEDULE)
    // multi-GPU exchange (comm.cu): one NCCL communicator per ctx, staging for the interleaved-band all-gather
    void *nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    void *comm_stage = nullptr;
    size_t comm_stage_bytes = 0;
    // peer-to-peer band exchange (comm.cu): every rank's staging block is mapped into every other rank (CUDA IPC), a rank
    // stores its bands straight into its peers' blocks over NVLink and raises a flag there
    static constexpr int P2P_MAX_RANKS = 16;
    static constexpr size_t P2P_HEADER_BYTES = 4096;  // flags[P2P_MAX_RANKS] at the start of the block, data behind
    int p2p_state = 0;             // 0: not tried yet, 1: in use, -1: unavailable (NCCL all-gather instead)
    void *p2p_block = nullptr;     // this rank's block: header + 2 x world x chunk (double-buffered by frame parity)
    size_t p2p_chunk_bytes = 0;
    void *p2p_peer[P2P_MAX_RANKS] = {};
    unsigned long long *p2p_counter = nullptr;
    uint32_t p2p_seq = 0;
    unsigned long long p2p_ctas = 0;  // CTAs of all pushes so far (what p2p_counter reaches when the latest push is complete)
    int *p2p_err_host = nullptr, *p2p_err_dev = nullptr;  // mapped page-locked: a wait that timed out
    int refs = 1;  // the ctx handle itself + every live scene / target: resources are freed when the last one goes
};

struct solb_scene {
    solb_ctx *ctx = nullptr;
    std::vector<SolbSceneInstance> instances;  // as the reference's shader would see them
    std::vector<DeviceInstance> h_inst;
    std::vector<uint32_t> h_first_tri;       // per-instance triangle prefix (flattened build)
    std::vector<DeviceBlas> h_blas;          // unique geometry: one per primitive section
    std::vector<SolbMaterialInfo> materials; // kept for solb_scene_add_instance
    uint32_t n_tris = 0, n_geom_tris = 0, n_vertices = 0, n_indices = 0;
    float4 *d_vertices = nullptr;
    uint32_t *d_indices = nullptr, *d_first_tri = nullptr;
    DeviceInstance *d_inst = nullptr;
    DeviceBlas *d_blas = nullptr;
    ShadeRecord *d_shade = nullptr;
    // base-colour textures (solb_scene_set_textures): linear float4 texels, one allocation per texture
    std::vector<float4 *> d_texels;
    TexDesc *d_tex = nullptr;
    uint32_t n_tex = 0;
    std::vector<uint32_t> material_texture;  // texture index per material (SOLB_NO_TEXTURE: none)
    TexBinding texb() const {
        TexBinding b;
        b.tex = d_tex;
        b.vertices = d_vertices;
        b.indices = d_indices;
        b.n_tex = n_tex;
        return b;
    }
    size_t inst_capacity = 0;                // instances d_inst / d_first_tri can hold
    AccelStorage accel;
    uint32_t accel_mode = SOLB_ACCEL_FLAT;
    bool built = false;
    bool dirty = false;       // an instance transform changed since the last build / TLAS regenerate
    bool needs_build = false; // instances were added or the mode changed: the whole structure must be rebuilt
    DeviceSceneView view() const {
        DeviceSceneView v;
        v.n_instances = (uint32_t)h_inst.size();
        v.n_tris = n_tris;
        v.inst_first_tri = d_first_tri;
        v.instances = d_inst;
        v.vertices = d_vertices;
        v.indices = d_indices;
        v.n_blas = (uint32_t)h_blas.size();
        v.n_geom_tris = n_geom_tris;
        v.blas = d_blas;
        return v;
    }
};

struct solb_target {
    solb_ctx *ctx = nullptr;
    uint32_t width = 0, height = 0, format = 0;
    void *dev = nullptr;
    size_t bytes = 0;
};

// hi_begin: *out = the stream to launch on, ordered after everything enqueued on the ctx stream so far; hi_end: the ctx stream
// continues after what was launched in between.  (use_hi_stream = 0: the ctx stream itself.)
static inline cudaError_t hi_begin(solb_ctx *c, cudaStream_t *out) {
    *out = c->stream;
    if (!c->use_hi_stream) return cudaSuccess;
    cudaError_t e;
    if (!c->hi_stream) {
        int least = 0, greatest = 0;
        if ((e = cudaDeviceGetStreamPriorityRange(&least, &greatest)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithPriority(&c->hi_stream, cudaStreamNonBlocking, greatest)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c->ev_hi_in, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&c->ev_hi_out, cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if ((e = cudaEventRecord(c->ev_hi_in, c->stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(c->hi_stream, c->ev_hi_in, 0)) != cudaSuccess) return e;
    *out = c->hi_stream;
    return cudaSuccess;
}
static inline cudaError_t hi_end(solb_ctx *c) {
    if (!c->use_hi_stream || !c->hi_stream) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaEventRecord(c->ev_hi_out, c->hi_stream)) != cudaSuccess) return e;
    return cudaStreamWaitEvent(c->stream, c->ev_hi_out, 0);
}

struct solb_fence {
    solb_ctx *ctx = nullptr;
    cudaEvent_t ev = nullptr;
    bool armed = false;  // false: created signalled, nothing to wait for
};

static inline int fail(solb_ctx *ctx, int code, const std::string &msg) {
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}
static inline int fail_cuda(solb_ctx *ctx, cudaError_t e, const char *what) {
    cudaGetLastError();  // clear sticky non-fatal state
    return fail(ctx, SOLB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define SOLB_TRY try {
#define SOLB_CATCH(ctx)                                                                  \
    } catch (const std::bad_alloc &) { return fail(ctx, SOLB_ERR_CUDA, "host out of memory"); } \
    catch (const std::exception &e) { return fail(ctx, SOLB_ERR_INVALID, e.what()); }       \
    catch (...) { return fail(ctx, SOLB_ERR_INVALID, "unknown exception"); }
#define CU(ctx, call)                                                    \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call);       \
    } while (0)

