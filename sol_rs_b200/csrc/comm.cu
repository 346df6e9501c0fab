// comm.cu — the path's multi-GPU exchange behind the C ABI (new work, the reference is single-GPU).
//
// One process per GPU, one NCCL communicator per solb_ctx.  Two exchanges exist on this path:
//   * frames split (sample split): every rank renders frames f = rank (mod N) into a per-rank SUM target; ONE float32
//     sum-reduce of the W x H x float4 buffers onto the root over NVLink, then the resolve (sum / count, gamma, rgba8) on the
//     root in the same call, on the ctx stream (solb_reduce_accum);
//   * tile split (single-sample interactive frames): rank r traces the 8-row bands r, r + N, ... of a frame; its rows are
//     packed, all-gathered and scattered back into every rank's full-size target (solb_allgather_rows).
// NCCL is bound at run time (dlopen): a host that already carries a copy (PyTorch bundles its own libnccl.so.2) must not get a
// second one mapped next to it, so the copy already in the process is used when there is one, else SOLB_NCCL_LIB, else the
// system library.  Nothing here runs without a communicator: no host-staged fallback.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; every call goes through the table below

#include "solb_handles.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char *names[3] = { nullptr, getenv("SOLB_NCCL_LIB"), "libnccl.so.2" };
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host application already loaded
    for (int i = 1; i < 3 && !h; i++)
        if (names[i] && *names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        api.error = std::string("NCCL not found (libnccl.so.2; set SOLB_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
        return &api;
    }
    api.handle = h;
#define SOLB_SYM(field, name)                                             \
    api.field = (decltype(api.field))dlsym(h, name);                      \
    if (!api.field) { api.error = std::string("NCCL symbol missing: ") + name; api.handle = nullptr; return &api; }
    SOLB_SYM(GetUniqueId, "ncclGetUniqueId")
    SOLB_SYM(CommInitRank, "ncclCommInitRank")
    SOLB_SYM(CommDestroy, "ncclCommDestroy")
    SOLB_SYM(Reduce, "ncclReduce")
    SOLB_SYM(AllGather, "ncclAllGather")
    SOLB_SYM(GetErrorString, "ncclGetErrorString")
    SOLB_SYM(GetVersion, "ncclGetVersion")
#undef SOLB_SYM
    return &api;
}

int fail_nccl(solb_ctx *ctx, ncclResult_t r, const char *what) {
    NcclApi *n = nccl_api();
    return fail(ctx, SOLB_ERR_CUDA, std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(r) : "NCCL error"));
}
#define NC(ctx, call)                                              \
    do {                                                           \
        ncclResult_t r__ = (call);                                 \
        if (r__ != ncclSuccess) return fail_nccl(ctx, r__, #call); \
    } while (0)

// Rows of the interleaved bands owned by `owner` <-> a contiguous chunk.  One thread per 16-byte (or smaller) element.
//   band b (rows [b * band_rows, (b + 1) * band_rows), clipped to height) belongs to rank b % world; the chunk holds the
//   owner's bands in ascending order, band_rows rows each (the last band of the image may be short: its tail is padding).
template <bool PACK>
__global__ void __launch_bounds__(256) k_band_rows(uint4 *image, uint4 *chunks, uint32_t row_elems, uint32_t height, uint32_t band_rows,
                                                   uint32_t world, uint32_t bands_per_rank, uint32_t only_owner, uint32_t skip_owner) {
    const uint32_t rows_per_chunk = bands_per_rank * band_rows;
    const uint64_t total = (uint64_t)(PACK ? 1u : world) * rows_per_chunk * row_elems;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % row_elems);
        const uint32_t cr = (uint32_t)(i / row_elems);
        const uint32_t owner = PACK ? only_owner : cr / rows_per_chunk;
        if (!PACK && owner == skip_owner) continue;  // this rank's own rows are already in place
        const uint32_t r = cr - (PACK ? 0u : owner * rows_per_chunk);
        const uint32_t band = (r / band_rows) * world + owner, y = band * band_rows + r % band_rows;
        if (y >= height) continue;
        const size_t ci = ((size_t)owner * rows_per_chunk + r) * row_elems + x, ii = (size_t)y * row_elems + x;
        if (PACK) chunks[ci] = image[ii];
        else image[ii] = chunks[ci];
    }
}

// ---- peer-to-peer exchange ------------------------------------------------------------------------
// The NCCL all-gather is a kernel of a few large CTAs; with frames in flight the persistent trace kernels of the next frames
// hold every SM, and a 512-thread CTA finds room only when a frame drains (measured at 8 GPUs: 3.19 ms per tile-split frame
// traced alone, 3.66 ms with the all-gather of the 8 MB rgba8 frame behind it, profiles/r02_tile_split.txt).  Here the
// exchange is made of 64-thread CTAs that fit wherever a trace warp has just left: k_p2p_push reads this rank's bands and
// stores them into every peer's staging block through NVLink (the blocks are mapped with CUDA IPC), the last CTA raises this
// rank's flag in every peer; k_p2p_wait (one warp) holds the stream until every peer's flag has arrived, and k_band_rows
// scatters the staged rows.  Blocks are double-buffered by sequence parity: a peer one exchange ahead writes the other half.
struct P2PPeers {
    uint4 *data[solb_ctx::P2P_MAX_RANKS];
    uint32_t *flags[solb_ctx::P2P_MAX_RANKS];
};

__global__ void __launch_bounds__(64) k_p2p_push(const uint4 *__restrict__ image, const P2PPeers peers, uint32_t row_elems, uint32_t height,
                                                 uint32_t band_rows, uint32_t world, uint32_t bands_per_rank, uint32_t me, uint32_t seq,
                                                 unsigned long long *counter, unsigned long long all_done) {
    const uint32_t rows_per_chunk = bands_per_rank * band_rows;
    const uint64_t total = (uint64_t)rows_per_chunk * row_elems;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % row_elems), r = (uint32_t)(i / row_elems);
        const uint32_t y = ((r / band_rows) * world + me) * band_rows + r % band_rows;
        const uint4 v = y < height ? image[(size_t)y * row_elems + x] : make_uint4(0u, 0u, 0u, 0u);
        const size_t ci = ((size_t)me * rows_per_chunk + r) * row_elems + x;
        for (uint32_t p = 0; p < world; p++)
            if (p != me) peers.data[p][ci] = v;
    }
    __threadfence_system();  // this thread's stores are ordered before whatever it (or a thread it synchronises with) stores next
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(counter, 1ull) + 1ull == all_done) {  // counter runs on: all_done = CTAs of all pushes so far
        __threadfence_system();
        for (uint32_t p = 0; p < world; p++)
            if (p != me) *(volatile uint32_t *)(peers.flags[p] + me) = seq;
    }
}

__global__ void __launch_bounds__(32) k_p2p_wait(const uint32_t *flags, uint32_t world, uint32_t me, uint32_t seq, int *err) {
    const uint32_t r = threadIdx.x;
    if (r < world && r != me) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int32_t)(*(volatile const uint32_t *)(flags + r) - seq) < 0) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 600000000000ull) { *(volatile int *)err = 1; break; }  // 10 min: a peer died; reported by the next call
        }
    }
    __threadfence_system();
}

// Collective: (re)allocate this rank's block for `chunk_bytes` per rank and parity, exchange the IPC handles, map the peers.
// Leaves p2p_state = 1 on every rank or -1 on every rank.
int p2p_setup(solb_ctx *ctx, size_t chunk_bytes) {
    NcclApi *n = nccl_api();
    const int world = ctx->comm_world, me = ctx->comm_rank;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    // nobody may still be storing into a block that is about to go: every rank drains its own stream, then all meet
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->hi_stream) CU(ctx, cudaStreamSynchronize(ctx->hi_stream));
    struct Slot { cudaIpcMemHandle_t handle; int ok; int pad[15]; };
    static_assert(sizeof(Slot) == 128, "Slot");
    Slot *d_slots = nullptr;
    std::vector<Slot> h_slots(world);
    CU(ctx, cudaMalloc((void **)&d_slots, sizeof(Slot) * world));
    auto gather = [&](const Slot &mine) -> int {
        CU(ctx, cudaMemcpyAsync(d_slots + me, &mine, sizeof(Slot), cudaMemcpyHostToDevice, ctx->stream));
        NC(ctx, n->AllGather(d_slots + me, d_slots, sizeof(Slot), ncclChar, comm, ctx->stream));
        CU(ctx, cudaMemcpyAsync(h_slots.data(), d_slots, sizeof(Slot) * world, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return SOLB_OK;
    };
    for (int p = 0; p < world; p++)
        if (ctx->p2p_peer[p]) { cudaIpcCloseMemHandle(ctx->p2p_peer[p]); ctx->p2p_peer[p] = nullptr; }
    Slot mine = {};
    int rc = gather(mine);  // the meeting point: from here on nobody maps (or stores into) anybody's old block
    if (rc != SOLB_OK) { cudaFree(d_slots); return rc; }
    cudaFree(ctx->p2p_block);
    ctx->p2p_block = nullptr;
    ctx->p2p_chunk_bytes = 0;
    ctx->p2p_seq = 0;
    ctx->p2p_ctas = 0;
    const size_t bytes = solb_ctx::P2P_HEADER_BYTES + 2 * (size_t)world * chunk_bytes;
    bool ok = world <= solb_ctx::P2P_MAX_RANKS && cudaMalloc(&ctx->p2p_block, bytes) == cudaSuccess &&
              cudaMemsetAsync(ctx->p2p_block, 0, bytes, ctx->stream) == cudaSuccess && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (ok && !ctx->p2p_counter) ok = cudaMalloc((void **)&ctx->p2p_counter, sizeof(unsigned long long)) == cudaSuccess;
    if (ok) ok = cudaMemsetAsync(ctx->p2p_counter, 0, sizeof(unsigned long long), ctx->stream) == cudaSuccess;
    if (ok && !ctx->p2p_err_host) {
        ok = cudaHostAlloc((void **)&ctx->p2p_err_host, sizeof(int), cudaHostAllocMapped) == cudaSuccess;
        if (ok) { *ctx->p2p_err_host = 0; ok = cudaHostGetDevicePointer((void **)&ctx->p2p_err_dev, ctx->p2p_err_host, 0) == cudaSuccess; }
    }
    if (ok) ok = cudaIpcGetMemHandle(&mine.handle, ctx->p2p_block) == cudaSuccess;
    cudaGetLastError();
    mine.ok = ok ? 1 : 0;
    rc = gather(mine);
    if (rc != SOLB_OK) { cudaFree(d_slots); return rc; }
    bool all = true;
    for (int p = 0; p < world; p++) all = all && h_slots[p].ok;
    if (all) {
        for (int p = 0; p < world && ok; p++) {
            if (p == me) continue;
            ok = cudaIpcOpenMemHandle(&ctx->p2p_peer[p], h_slots[p].handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (!ok) ctx->p2p_peer[p] = nullptr;
        }
        cudaGetLastError();
    }
    mine.ok = (all && ok) ? 1 : 0;
    rc = gather(mine);  // every rank learns whether every rank mapped every block
    cudaFree(d_slots);
    if (rc != SOLB_OK) return rc;
    for (int p = 0; p < world; p++) all = all && h_slots[p].ok;
    if (all) {
        ctx->p2p_state = 1;
        ctx->p2p_chunk_bytes = chunk_bytes;
    } else {
        for (int p = 0; p < world; p++)
            if (ctx->p2p_peer[p]) { cudaIpcCloseMemHandle(ctx->p2p_peer[p]); ctx->p2p_peer[p] = nullptr; }
        cudaFree(ctx->p2p_block);
        ctx->p2p_block = nullptr;
        ctx->p2p_state = -1;
    }
    return SOLB_OK;
}

void p2p_release(solb_ctx *ctx) {
    for (int p = 0; p < solb_ctx::P2P_MAX_RANKS; p++)
        if (ctx->p2p_peer[p]) { cudaIpcCloseMemHandle(ctx->p2p_peer[p]); ctx->p2p_peer[p] = nullptr; }
    cudaFree(ctx->p2p_block);
    cudaFree(ctx->p2p_counter);
    if (ctx->p2p_err_host) cudaFreeHost(ctx->p2p_err_host);
    ctx->p2p_block = nullptr;
    ctx->p2p_counter = nullptr;
    ctx->p2p_err_host = ctx->p2p_err_dev = nullptr;
    ctx->p2p_chunk_bytes = 0;
    ctx->p2p_state = 0;
    ctx->p2p_seq = 0;
    ctx->p2p_ctas = 0;
}

}  // namespace

extern "C" {

SOLB_API int solb_comm_unique_id(uint8_t *id_out) {
    if (!id_out) return fail(nullptr, SOLB_ERR_INVALID, "solb_comm_unique_id: null argument");
    NcclApi *n = nccl_api();
    if (!n->handle) return fail(nullptr, SOLB_ERR_UNSUPPORTED, n->error);
    static_assert(sizeof(ncclUniqueId) == SOLB_COMM_ID_BYTES, "SOLB_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
    ncclUniqueId id;
    NC(nullptr, n->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return SOLB_OK;
}

SOLB_API int solb_comm_init(solb_ctx *ctx, const uint8_t *id_bytes, int rank, int world) {
    if (!ctx || !id_bytes) return fail(ctx, SOLB_ERR_INVALID, "solb_comm_init: null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, SOLB_ERR_INVALID, "solb_comm_init: bad rank / world");
    if (ctx->nccl_comm) return fail(ctx, SOLB_ERR_INVALID, "solb_comm_init: this ctx already has a communicator");
    NcclApi *n = nccl_api();
    if (!n->handle) return fail(ctx, SOLB_ERR_UNSUPPORTED, n->error);
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    NC(ctx, n->CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    // NCCL connects the ranks lazily, at the first collective of each kind (measured: 180 - 410 ms inside the first
    // ncclReduce of a fresh communicator).  Pay for it here, where set-up belongs, with one tiny reduce and all-gather.
    {
        float *scratch = nullptr;
        CU(ctx, cudaMalloc((void **)&scratch, sizeof(float) * 4 * (size_t)world));
        CU(ctx, cudaMemsetAsync(scratch, 0, sizeof(float) * 4 * (size_t)world, ctx->stream));
        ncclResult_t r = n->Reduce(scratch, scratch, 4, ncclFloat32, ncclSum, 0, comm, ctx->stream);
        if (r == ncclSuccess) r = n->AllGather(scratch + 4 * rank, scratch, 4, ncclFloat32, comm, ctx->stream);
        const cudaError_t e = cudaStreamSynchronize(ctx->stream);
        cudaFree(scratch);
        if (r != ncclSuccess) return fail_nccl(ctx, r, "communicator warm-up");
        if (e != cudaSuccess) return fail_cuda(ctx, e, "communicator warm-up");
    }
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

SOLB_API int solb_comm_info(solb_ctx *ctx, int *rank, int *world, int *nccl_version) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->nccl_comm ? ctx->comm_world : 1;
    if (nccl_version) {
        *nccl_version = 0;
        NcclApi *n = nccl_api();
        if (n->handle) n->GetVersion(nccl_version);
    }
    return SOLB_OK;
}

SOLB_API int solb_comm_destroy(solb_ctx *ctx) {
    if (!ctx) return SOLB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->hi_stream) cudaStreamSynchronize(ctx->hi_stream);
    if (ctx->nccl_comm) {
        // (a peer stores into this rank's block only inside an exchange this rank takes part in, and this rank's wait kernel of
        //  that exchange - drained above - returns after every peer's stores: nothing is in flight towards the block now)
        p2p_release(ctx);
        nccl_api()->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    cudaFree(ctx->comm_stage);
    ctx->comm_stage = nullptr;
    ctx->comm_stage_bytes = 0;
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    return SOLB_OK;
}

// Frames split: sum the per-rank SOLB_ACCUM_SUM targets onto `root` (in place) and resolve there, all on the ctx stream.
SOLB_API int solb_reduce_accum(solb_ctx *ctx, solb_target *sum, int root, solb_target *accum_out, solb_target *render) {
    if (!ctx || !sum) return fail(ctx, SOLB_ERR_INVALID, "solb_reduce_accum: null argument");
    if (sum->ctx != ctx || sum->format != SOLB_FORMAT_RGBA32F) return fail(ctx, SOLB_ERR_INVALID, "solb_reduce_accum: sum must be an rgba32f target of this ctx");
    const int world = ctx->nccl_comm ? ctx->comm_world : 1;
    if (root < 0 || root >= world) return fail(ctx, SOLB_ERR_INVALID, "solb_reduce_accum: bad root");
    if (world > 1 && !ctx->nccl_comm) return fail(ctx, SOLB_ERR_INVALID, "solb_reduce_accum before solb_comm_init");
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    if (world > 1) {
        NC(ctx, nccl_api()->Reduce(sum->dev, sum->dev, (size_t)sum->width * sum->height * 4, ncclFloat32, ncclSum, root,
                                   (ncclComm_t)ctx->nccl_comm, ctx->stream));
        ctx->launches += 1;
    }
    if (ctx->comm_rank == root && (accum_out || render)) return solb_resolve_sum(ctx, sum, accum_out, render);
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

// Tile split: every rank has traced its interleaved bands (SolbTraceParams tile_row_begin = rank * band_rows, tile_row_count =
// band_rows, tile_row_stride = world * band_rows) into `target`; afterwards every rank holds the whole image.
SOLB_API int solb_allgather_rows(solb_ctx *ctx, solb_target *target, uint32_t band_rows) {
    if (!ctx || !target) return fail(ctx, SOLB_ERR_INVALID, "solb_allgather_rows: null argument");
    if (target->ctx != ctx) return fail(ctx, SOLB_ERR_INVALID, "solb_allgather_rows: target belongs to another ctx");
    if (band_rows == 0) return fail(ctx, SOLB_ERR_INVALID, "solb_allgather_rows: band_rows must be > 0");
    const int world = ctx->nccl_comm ? ctx->comm_world : 1;
    if (world == 1) return SOLB_OK;
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t bpp = target->format == SOLB_FORMAT_RGBA32F ? 16 : (target->format == SOLB_FORMAT_RGBA8 ? 4 : 8);
    const size_t row_bytes = (size_t)target->width * bpp;
    if (row_bytes % 16) return fail(ctx, SOLB_ERR_UNSUPPORTED, "solb_allgather_rows: row size must be a multiple of 16 bytes");
    const uint32_t n_bands = (target->height + band_rows - 1) / band_rows;
    const uint32_t bands_per_rank = (n_bands + world - 1) / world;
    const size_t chunk_bytes = (size_t)bands_per_rank * band_rows * row_bytes;
    const uint32_t row_elems = (uint32_t)(row_bytes / 16);
    if (ctx->p2p_err_host && *ctx->p2p_err_host) return fail(ctx, SOLB_ERR_CUDA, "solb_allgather_rows: a peer did not deliver its bands within 10 min");
    static const bool want_p2p = !(getenv("SOLB_P2P") && atoi(getenv("SOLB_P2P")) == 0);
    if (want_p2p && (ctx->p2p_state == 0 || (ctx->p2p_state == 1 && ctx->p2p_chunk_bytes < chunk_bytes))) {
        // sized for 16-byte texels at once: the rgba8 frame and the float accumulation of one image alternate through one block
        const int rc = p2p_setup(ctx, std::max(chunk_bytes, (size_t)bands_per_rank * band_rows * target->width * 16));
        if (rc != SOLB_OK) return rc;
    }
    if (want_p2p && ctx->p2p_state == 1) {
        const uint32_t seq = ++ctx->p2p_seq;
        const size_t half = (size_t)world * ctx->p2p_chunk_bytes, off = solb_ctx::P2P_HEADER_BYTES + (seq & 1u) * half;
        P2PPeers peers = {};
        for (int p = 0; p < world; p++) {
            char *base = (char *)(p == ctx->comm_rank ? ctx->p2p_block : ctx->p2p_peer[p]);
            peers.data[p] = (uint4 *)(base + off);
            peers.flags[p] = (uint32_t *)base;
        }
        const uint64_t total = (uint64_t)bands_per_rank * band_rows * row_elems;
        const uint32_t push_grid = (uint32_t)std::min<uint64_t>((total + 63) / 64, (uint64_t)ctx->sm_count * 16);
        cudaStream_t xs;
        CU(ctx, hi_begin(ctx, &xs));
        k_p2p_push<<<push_grid, 64, 0, xs>>>((const uint4 *)target->dev, peers, row_elems, target->height, band_rows, (uint32_t)world,
                                            bands_per_rank, (uint32_t)ctx->comm_rank, seq, ctx->p2p_counter,
                                            ctx->p2p_ctas += push_grid);
        k_p2p_wait<<<1, 32, 0, xs>>>((const uint32_t *)ctx->p2p_block, (uint32_t)world, (uint32_t)ctx->comm_rank, seq, ctx->p2p_err_dev);
        k_band_rows<false><<<ctx->sm_count * 16, 64, 0, xs>>>((uint4 *)target->dev, (uint4 *)((char *)ctx->p2p_block + off), row_elems, target->height,
                                                             band_rows, (uint32_t)world, bands_per_rank, 0u, (uint32_t)ctx->comm_rank);
        CU(ctx, cudaGetLastError());
        CU(ctx, hi_end(ctx));
        ctx->launches += 3;
        return SOLB_OK;
    }
    if (ctx->comm_stage_bytes < chunk_bytes * world) {
        cudaFree(ctx->comm_stage);
        ctx->comm_stage = nullptr;
        ctx->comm_stage_bytes = 0;
        CU(ctx, cudaMalloc(&ctx->comm_stage, chunk_bytes * world));
        CU(ctx, cudaMemsetAsync(ctx->comm_stage, 0, chunk_bytes * world, ctx->stream));  // padding rows are sent too
        ctx->comm_stage_bytes = chunk_bytes * world;
    }
    const int grid = ctx->sm_count * 4;
    cudaStream_t xs;  // ordered inside the ctx stream, but ahead of the trace kernels of the next frames in the queue for SM space
    CU(ctx, hi_begin(ctx, &xs));
    k_band_rows<true><<<grid, 256, 0, xs>>>((uint4 *)target->dev, (uint4 *)ctx->comm_stage, row_elems, target->height, band_rows,
                                                    (uint32_t)world, bands_per_rank, (uint32_t)ctx->comm_rank, 0u);
    CU(ctx, cudaGetLastError());
    NC(ctx, nccl_api()->AllGather((const char *)ctx->comm_stage + chunk_bytes * ctx->comm_rank, ctx->comm_stage, chunk_bytes, ncclChar,
                                  (ncclComm_t)ctx->nccl_comm, xs));
    k_band_rows<false><<<grid, 256, 0, xs>>>((uint4 *)target->dev, (uint4 *)ctx->comm_stage, row_elems, target->height, band_rows,
                                                     (uint32_t)world, bands_per_rank, 0u, (uint32_t)ctx->comm_rank);
    CU(ctx, cudaGetLastError());
    CU(ctx, hi_end(ctx));
    ctx->launches += 3;
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

}  // extern "C"
