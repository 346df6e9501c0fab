// trace.cu — trace kernels for sm_100a: the three reference pipelines as CUDA launches.
//   5-pathtrace  (assets/glsl/pathtrace.{rgen,rchit,rmiss})  -> wavefront kernels or megakernel
//   4-ray-ao     (assets/glsl/ao.{rgen,rchit,rmiss})         -> k_ao
//   3-ray-debug  (assets/glsl/debug.{rgen,rchit,rmiss})      -> k_debug
// Each launch sequence replaces one ShaderBindingTable::cmd_trace_rays (src/ray/sbt.rs:167-180).
#include "trace.h"

#include <algorithm>

#include "shade.cuh"

namespace solb {

constexpr int TRACE_BLOCK = 128;
// threads per CTA of the warp-local wavefront kernel.  Its warps never talk to each other, so the CTA is only the unit in
// which the SM hands out and takes back warp slots: with frames in flight a CTA of the next frame starts when ALL warps of a
// CTA of the draining frame have finished.  One warp per CTA: 4 516 Mrays/s on the headline against 4 415 with two and 4 403 with
// four warps per CTA, and 4 221 against 4 148 without frames in flight (profiles/r02_frames_in_flight.txt).
#ifndef SOLB_WL_BLOCK
#define SOLB_WL_BLOCK 32
#endif
constexpr int WL_BLOCK = SOLB_WL_BLOCK;
static_assert(WL_BLOCK % 32 == 0 && WL_BLOCK >= 32 && WL_BLOCK <= TRACE_BLOCK && TRACE_BLOCK % WL_BLOCK == 0, "SOLB_WL_BLOCK");

// Traversal stack: SOLB_SM_STACK entries per lane in shared memory ([entry][thread] so a warp's
// accesses are conflict-free), the rest spills to local memory.
// The shared part is addressed in the shared state space (32-bit address, st.shared / ld.shared): through a generic pointer
// every push and pop rebuilt the window base from special registers (S2R tid, S2R cga id, MOV, 2 x LEA) - 12 to 17
// instructions per stack operation, three operations per iteration of the traversal loop
// (profiles/r02_ncu_k_pt_warpfront_d.txt).
template <int BLOCK>
struct DevStackT {
    uint32_t sm;  // shared-space byte address of this lane's column; entry e at sm + e * BLOCK * 8
    uint2 *loc;   // spill array (local memory); kept OUTSIDE the struct so sp / sm stay in registers
    int sp;
    __device__ __forceinline__ void push(uint2 v) {
        if (sp < SOLB_SM_STACK) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sm + (uint32_t)sp * (BLOCK * 8u)), "r"(v.x), "r"(v.y) : "memory");
        else loc[sp - SOLB_SM_STACK] = v;
        sp++;
    }
    __device__ __forceinline__ uint2 pop() {
        sp--;
        uint2 v;
        if (sp < SOLB_SM_STACK) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sm + (uint32_t)sp * (BLOCK * 8u)) : "memory");
        else v = loc[sp - SOLB_SM_STACK];
        return v;
    }
    __device__ __forceinline__ bool empty() const { return sp == 0; }
};
using DevStack = DevStackT<TRACE_BLOCK>;

#define SOLB_DECL_STACK() SOLB_DECL_STACK_N(TRACE_BLOCK)
#define SOLB_DECL_STACK_N(BLOCK)                                               \
    __shared__ uint2 s_stack[SOLB_SM_STACK * (BLOCK)];                         \
    uint2 stack_spill[SOLB_LOCAL_STACK];                                       \
    DevStackT<(BLOCK)> stack;                                                  \
    stack.sm = (uint32_t)__cvta_generic_to_shared(s_stack + threadIdx.x);      \
    stack.loc = stack_spill;                                                   \
    stack.sp = 0

// flattened (TL = false) or two-level (TL = true) closest hit
template <bool STATS, bool TL, class Stack>
__device__ __forceinline__ void trace_any(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris,
                                          const float4 *__restrict__ inst_leaves, const Ray &ray, Hit &hit, Stack &stack,
                                          TraceCounters *ctr) {
    if (TL) trace_closest_2l<STATS>(nodes, tris, inst_leaves, ray, hit, stack, ctr);
    else trace_closest<STATS>(nodes, tris, ray, hit, stack, ctr);
}

// Warp-voted closest hit for the one-pixel-per-lane kernels (megakernel, AO, debug).  trace_closest lets every lane run its
// own node-step / triangle-loop sequence, so the warp executes the union of all lanes' sequences: in the megakernel's first
// profile the triangle loop ran 3.7 lanes wide and the node step 10.7 (profiles/r01_ncu_k_pathtrace_mega_b.txt).  Here the
// whole warp steps together like k_wf_trace does: every iteration a ballot picks the step more lanes are waiting for, lanes
// waiting for the other kind postpone (node step: pending triangles go on the stack), and lanes whose ray is finished idle
// until the last one is.  Must be called by all 32 lanes of the warp (has_ray = false for lanes without a ray).  Visiting order
// per ray differs from trace_closest only in WHEN postponed triangle groups are tested, not in which.
template <bool STATS, bool TL, class Stack>
__device__ __forceinline__ void trace_vote(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris,
                                           const float4 *__restrict__ inst_leaves, const Ray &ray, bool has_ray, Hit &hit,
                                           Stack &stack, TraceCounters *ctr) {
    hit.inst = SOLB_MISS; hit.prim = SOLB_MISS; hit.gtri = SOLB_MISS;
    hit.t = ray.tmax; hit.u = 0.0f; hit.v = 0.0f;
    TravRay tr = make_trav_ray(ray.o, ray.d, ray.tmin);
    float tmax = ray.tmax;
    uint2 ngroup = has_ray ? SOLB_ROOT_GROUP : make_uint2(0u, 0u);
    uint2 tgroup = make_uint2(0u, 0u);
    bool in_blas = false;           // TL only
    uint32_t cur_inst = SOLB_MISS;  // TL only
    stack.sp = 0;
    for (;;) {
        const bool w_node = (ngroup.y & 0xff000000u) != 0u, w_tri = tgroup.y != 0u;
        const uint32_t b_node = __ballot_sync(0xffffffffu, w_node), b_tri = __ballot_sync(0xffffffffu, w_tri);
        // (a lane that has just popped a TLAS sentinel holds nothing but may still have entries on its stack)
        if ((b_node | b_tri) == 0u && !__any_sync(0xffffffffu, !stack.empty())) break;
        if (b_tri != 0u && __popc(b_tri) >= __popc(b_node)) {
            if (w_tri) {
                if (TL && !in_blas) {
                    trav_enter_instance(inst_leaves, ray.o, ray.d, tr, ngroup, tgroup, cur_inst, stack);
                    in_blas = true;
                } else {
                    if (trav_tri_step(tris, tr, tmax, tgroup, hit) && TL) hit.inst = cur_inst;
                    if (STATS) ctr->tris++;
                }
            }
        } else if (w_node) {
            if (tgroup.y) stack.push(tgroup);  // postpone the pending triangles
            trav_node_step(nodes, tr, tmax, ngroup, tgroup, stack);
            if (STATS) ctr->nodes++;
        }
        if (!(ngroup.y & 0xff000000u) && !tgroup.y && !stack.empty()) {
            const uint2 e = stack.pop();
            if (TL && e.y == 0u) {  // sentinel: back to the TLAS with the world-space ray
                tr = make_trav_ray(ray.o, ray.d, ray.tmin);
                in_blas = false;
            } else if (e.y & 0xff000000u) ngroup = e;
            else tgroup = e;
        }
    }
}

// stats slots (unsigned long long each)
enum { ST_RAYS = 0, ST_HITS = 1, ST_PATHS = 2, ST_NODES = 3, ST_TRIS = 4 };

__device__ __forceinline__ void warp_add_stat(unsigned long long *stats, int slot, uint32_t v) {
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&stats[slot], (unsigned long long)v);
}

// pixel owned by this thread: each warp covers an 8x4 tile so primary rays stay coherent
// The launch covers fc.n_bands bands of image rows (the whole image, or one rank's share of a tile-split frame):
// pixel coordinates, RNG seeds and target addresses are always those of the full image.  Region-local 8x4 pixel tile
// (tx, ty) + lane -> pixel.
__host__ __device__ __forceinline__ uint32_t tiles_per_band(const FrameConsts &fc) { return (fc.band_rows + 3u) >> 2; }
__device__ __forceinline__ bool region_pixel(const FrameConsts &fc, uint32_t tx, uint32_t ty, uint32_t lane, uint32_t &x, uint32_t &y) {
    const uint32_t tpb = tiles_per_band(fc);
    const uint32_t band = ty / tpb, row_in_band = (ty - band * tpb) * 4u + (lane >> 3);
    x = tx * 8u + (lane & 7u);
    y = fc.row_begin + band * fc.band_stride + row_in_band;
    return x < fc.width && band < fc.n_bands && row_in_band < fc.band_rows && y < fc.height;
}
__device__ __forceinline__ bool thread_pixel(const FrameConsts &fc, uint32_t &x, uint32_t &y) {
    const uint32_t tiles_x = (fc.width + 7u) >> 3;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    const uint32_t ty = gwarp / tiles_x, tx = gwarp - ty * tiles_x;
    return region_pixel(fc, tx, ty, lane, x, y);
}
static uint32_t region_tiles_y(const FrameConsts &fc) { return tiles_per_band(fc) * fc.n_bands; }
static uint32_t pixel_grid_blocks(const FrameConsts &fc) {
    const uint64_t warps = (uint64_t)((fc.width + 7u) >> 3) * region_tiles_y(fc);
    return (uint32_t)((warps * 32u + TRACE_BLOCK - 1) / TRACE_BLOCK);
}

// ---------------------------------------------------------------------------------------------------
// 3-ray-debug: debug.rgen:18-37 + debug.rchit:9-13 + debug.rmiss:6-9, plus the ids the parity gate needs
template <bool TL, bool VOTE>
__global__ void __launch_bounds__(TRACE_BLOCK) k_debug(const FrameConsts fc, const uint4 *__restrict__ nodes,
                                                       const float4 *__restrict__ tris, const float4 *__restrict__ inst_leaves,
                                                       uint32_t *render, uint2 *ids, float4 *attribs, unsigned long long *stats) {
    SOLB_DECL_STACK();
    uint32_t x, y;
    const bool active = thread_pixel(fc, x, y);
    uint32_t nr = 0, nh = 0;
    Ray r;
    r.o = fc.origin;
    r.d = active ? primary_dir(fc, (float)x + 0.5f, (float)y + 0.5f) : f3(0, 0, 1);  // debug.rgen:20
    r.tmin = 0.001f;                                                                 // debug.rgen:32-33
    r.tmax = 1000.0f;
    Hit h;
    if (VOTE) trace_vote<false, TL>(nodes, tris, inst_leaves, r, active, h, stack, (TraceCounters *)nullptr);
    else if (active) trace_any<false, TL>(nodes, tris, inst_leaves, r, h, stack, (TraceCounters *)nullptr);
    if (active) {
        nr = 1;
        float3 hv = r.d;  // debug.rgen:28: payload preset to the direction, the miss shader leaves it
        if (h.inst != SOLB_MISS) { hv = f3(1.0f - h.u - h.v, h.u, h.v); nh = 1; }  // debug.rchit:11-12
        const size_t p = (size_t)y * fc.width + x;
        if (render) render[p] = pack_rgba8(hv.x, hv.y, hv.z, 0.0f);  // debug.rgen:36
        if (ids) ids[p] = make_uint2(h.inst, h.prim);
        if (attribs) attribs[p] = make_float4(h.u, h.v, h.inst != SOLB_MISS ? h.t : 0.0f, 0.0f);
    }
    warp_add_stat(stats, ST_RAYS, nr);
    warp_add_stat(stats, ST_HITS, nh);
}

// traceRayEXT for arbitrary rays (tests)
template <bool TL, bool VOTE>
__global__ void __launch_bounds__(TRACE_BLOCK) k_trace_rays(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris,
                                                            const float4 *__restrict__ inst_leaves, const float4 *__restrict__ rays,
                                                            uint32_t n, uint4 *hits, float *t_out, unsigned long long *stats) {
    SOLB_DECL_STACK();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    TraceCounters ctr = { 0, 0 };
    uint32_t nr = 0;
    Ray r;
    r.o = f3(0, 0, 0); r.d = f3(0, 0, 1); r.tmin = 0.0f; r.tmax = 0.0f;
    if (i < n) {
        const float4 a = rays[2 * (size_t)i], b = rays[2 * (size_t)i + 1];
        r.o = f3(a.x, a.y, a.z); r.tmin = a.w;
        r.d = f3(b.x, b.y, b.z); r.tmax = b.w;
    }
    Hit h;
    if (VOTE) trace_vote<true, TL>(nodes, tris, inst_leaves, r, i < n, h, stack, &ctr);
    else if (i < n) trace_any<true, TL>(nodes, tris, inst_leaves, r, h, stack, &ctr);
    if (i < n) {
        hits[i] = make_uint4(h.inst, h.prim, __float_as_uint(h.u), __float_as_uint(h.v));
        if (t_out) t_out[i] = h.inst != SOLB_MISS ? h.t : 0.0f;
        nr = 1;
    }
    warp_add_stat(stats, ST_RAYS, nr);
    warp_add_stat(stats, ST_NODES, ctr.nodes);
    warp_add_stat(stats, ST_TRIS, ctr.tris);
}

// ---------------------------------------------------------------------------------------------------
// 5-pathtrace, megakernel schedule: pathtrace.rgen:39-104 with the sample and bounce loops flattened
// into one loop so a lane that ends a path immediately starts its next sample.
template <bool STATS, bool TL, bool VOTE>
__global__ void __launch_bounds__(TRACE_BLOCK, (STATS || TL) ? 1 : 6) k_pathtrace_mega(const FrameConsts fc, const uint4 *__restrict__ nodes,
                                                                const float4 *__restrict__ tris,
                                                                const float4 *__restrict__ inst_leaves,
                                                                const DeviceInstance *__restrict__ instances,
                                                                const ShadeRecord *__restrict__ shade, float4 *accum,
                                                                uint32_t *render, unsigned long long *stats) {
    SOLB_DECL_STACK();
    uint32_t x, y;
    const bool active = thread_pixel(fc, x, y);
    uint32_t nr = 0, nh = 0, np = 0;
    TraceCounters ctr = { 0, 0 };
    uint32_t rng = 0, sample = 0, depth = 0;
    float3 pixel = f3(0, 0, 0), thr = f3(1, 1, 1);
    Ray r;
    r.o = fc.origin; r.d = f3(0, 0, 1);
    r.tmin = fc.tmin;  // rayRange is set once per sample and never touched by rchit (:35)
    r.tmax = fc.tmax;
    bool live = active && fc.spp > 0u;
    if (live) {
        rng = tea(x + y * fc.width, fc.frame);                 // :47
        const float jx = next_rand(rng), jy = next_rand(rng);  // :52
        r.d = primary_dir(fc, (float)x + jx, (float)y + jy);
        np++;
    }
    // VOTE: the whole warp stays in the loop until its last pixel is done, so the traversal can step warp-wide
    while (VOTE ? __any_sync(0xffffffffu, live) : live) {
        Hit h;
        if (VOTE) {
            trace_vote<STATS, TL>(nodes, tris, inst_leaves, r, live, h, stack, &ctr);  // :65-76
        } else {
            stack.sp = 0;
            trace_any<STATS, TL>(nodes, tris, inst_leaves, r, h, stack, &ctr);
        }
        if (live) {
            nr++;
            bool end_path;
            if (h.inst != SOLB_MISS) {
                nh++;
                float3 hv;
                const bool done = shade_hit(instances, shade, fc.texb, h.inst, h.gtri, h.u, h.v, r.o, r.d, rng, hv);
                depth++;
                thr = thr * hv;  // :77
                end_path = done;
                if (!done && depth > fc.max_bounces) {  // :81-84
                    thr = f3(0, 0, 0);
                    end_path = true;
                }
            } else {
                thr = thr * shade_miss(fc.enable_sky, r.d);  // rmiss, done = 1
                end_path = true;
            }
            if (end_path) {
                pixel = pixel + thr;  // :86
                sample++;
                if (sample < fc.spp) {
                    const float jx = next_rand(rng), jy = next_rand(rng);
                    r.o = fc.origin;
                    r.d = primary_dir(fc, (float)x + jx, (float)y + jy);
                    thr = f3(1, 1, 1);
                    depth = 0;
                    np++;
                } else {
                    live = false;
                }
            }
        }
    }
    if (active) {
        const size_t p = (size_t)y * fc.width + x;
        uint32_t rgba;
        const float4 out = resolve_pixel(fc, pixel, accum[p], rgba);
        accum[p] = out;
        if (render) render[p] = rgba;
    }
    warp_add_stat(stats, ST_RAYS, nr);
    warp_add_stat(stats, ST_HITS, nh);
    warp_add_stat(stats, ST_PATHS, np);
    if (STATS) {
        warp_add_stat(stats, ST_NODES, ctr.nodes);
        warp_add_stat(stats, ST_TRIS, ctr.tris);
    }
}

__device__ __forceinline__ uint32_t swizzled_pixel(uint32_t i, const FrameConsts &fc, bool &valid);

// Persistent variant of the megakernel (experimental, SOLB_MEGA_PERSISTENT=1): a lane that finishes its pixel's samples
// resolves it and takes the next pixel from a global counter (8x4-tile order, warp-aggregated batches), so a warp no longer
// waits for its longest pixel (the per-pixel ray count varies 8..72 on the shipped scenes).  Per-pixel arithmetic and RNG
// streams are unchanged.  Measured 5 % slower than one thread per pixel (TraceTuning::mega_persistent): kept for comparison.
constexpr uint32_t MEGA_BATCH = 64;
template <bool STATS, bool TL, bool VOTE>
__global__ void __launch_bounds__(TRACE_BLOCK, (STATS || TL) ? 1 : 6) k_pathtrace_mega_persistent(const FrameConsts fc, const uint4 *__restrict__ nodes,
                                                                           const float4 *__restrict__ tris,
                                                                           const float4 *__restrict__ inst_leaves,
                                                                           const DeviceInstance *__restrict__ instances,
                                                                           const ShadeRecord *__restrict__ shade, float4 *accum,
                                                                           uint32_t *render, unsigned long long *stats,
                                                                           uint32_t *pixel_counter, uint32_t n_slots, int fetch_idle) {
    SOLB_DECL_STACK();
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    uint32_t nr = 0, nh = 0, np = 0;
    TraceCounters ctr = { 0, 0 };
    uint32_t pool_next = 0, pool_end = 0;  // warp-uniform range of tile-order slots
    bool exhausted = false, has_pixel = false;
    uint32_t x = 0, y = 0, rng = 0, sample = 0, depth = 0;
    float3 pixel = f3(0, 0, 0), thr = f3(1, 1, 1);
    Ray r;
    r.o = fc.origin; r.d = f3(0, 0, 1); r.tmin = fc.tmin; r.tmax = fc.tmax;
    uint32_t b_pix = 0;
    for (;;) {
        const uint32_t idle = ~b_pix;
        if (!exhausted && __popc(idle) >= fetch_idle) {
            const uint32_t want = (uint32_t)__popc(idle);
            uint32_t served = 0;
            while (served < want && !exhausted) {
                if (pool_next >= pool_end) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(pixel_counter, MEGA_BATCH);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base >= n_slots) { exhausted = true; break; }
                    pool_next = base;
                    pool_end = min(base + MEGA_BATCH, n_slots);
                }
                const uint32_t take = min(want - served, pool_end - pool_next);
                const uint32_t rank = (uint32_t)__popc(idle & lt_mask);
                if (!has_pixel && rank >= served && rank < served + take) {
                    bool valid;
                    const uint32_t p = swizzled_pixel(pool_next + (rank - served), fc, valid);
                    if (valid) {  // image edges leave holes in the tile order
                        x = p % fc.width; y = p / fc.width;
                        rng = tea(p, fc.frame);  // pathtrace.rgen:47
                        pixel = f3(0, 0, 0); thr = f3(1, 1, 1);
                        sample = 0; depth = 0;
                        const float jx = next_rand(rng), jy = next_rand(rng);  // :52
                        r.o = fc.origin;
                        r.d = primary_dir(fc, (float)x + jx, (float)y + jy);
                        np++;
                        has_pixel = true;
                    }
                }
                pool_next += take;
                served += take;
            }
            b_pix = __ballot_sync(0xffffffffu, has_pixel);
        }
        if (b_pix == 0u) {
            if (exhausted) break;
            continue;  // every slot of this round was an edge hole: fetch again
        }
        Hit h;
        if (VOTE) {
            trace_vote<STATS, TL>(nodes, tris, inst_leaves, r, has_pixel, h, stack, &ctr);  // :65-76
        } else if (has_pixel) {
            stack.sp = 0;
            trace_any<STATS, TL>(nodes, tris, inst_leaves, r, h, stack, &ctr);
        }
        if (has_pixel) {
            nr++;
            bool end_path;
            if (h.inst != SOLB_MISS) {
                nh++;
                float3 hv;
                const bool done = shade_hit(instances, shade, fc.texb, h.inst, h.gtri, h.u, h.v, r.o, r.d, rng, hv);
                depth++;
                thr = thr * hv;  // :77
                end_path = done;
                if (!done && depth > fc.max_bounces) {  // :81-84
                    thr = f3(0, 0, 0);
                    end_path = true;
                }
            } else {
                thr = thr * shade_miss(fc.enable_sky, r.d);  // rmiss, done = 1
                end_path = true;
            }
            if (end_path) {
                pixel = pixel + thr;  // :86
                sample++;
                if (sample < fc.spp) {
                    const float jx = next_rand(rng), jy = next_rand(rng);
                    r.o = fc.origin;
                    r.d = primary_dir(fc, (float)x + jx, (float)y + jy);
                    thr = f3(1, 1, 1);
                    depth = 0;
                    np++;
                } else {
                    const size_t p = (size_t)y * fc.width + x;
                    uint32_t rgba;
                    const float4 out = resolve_pixel(fc, pixel, accum[p], rgba);
                    accum[p] = out;
                    if (render) render[p] = rgba;
                    has_pixel = false;
                }
            }
        }
        b_pix = __ballot_sync(0xffffffffu, has_pixel);
    }
    warp_add_stat(stats, ST_RAYS, nr);
    warp_add_stat(stats, ST_HITS, nh);
    warp_add_stat(stats, ST_PATHS, np);
    if (STATS) {
        warp_add_stat(stats, ST_NODES, ctr.nodes);
        warp_add_stat(stats, ST_TRIS, ctr.tris);
    }
}

// ---------------------------------------------------------------------------------------------------
// 5-pathtrace, wavefront schedule.
//   k_wf_generate : per pixel, seed the RNG, emit sample 0's primary ray, fill the ray queue
//   k_wf_trace    : persistent warps pull 32-ray batches from the queue (atomic head), closest hit -> hit record
//   k_wf_shade    : per queued pixel, rchit / rmiss / bounce cap, regenerate the pixel's next sample in place,
//                   ballot-compact the still-alive pixels into the next queue
//   k_wf_resolve  : per pixel, mean of spp, running mix, NaN/Inf guard, gamma, rgba8
// Per-pixel path state lives in SoA float4/uint4 arrays (WavefrontState); the RNG state flows from sample to
// sample inside a pixel exactly like prd.rng does in pathtrace.rgen:47-86.

__device__ __forceinline__ uint32_t swizzled_pixel(uint32_t i, const FrameConsts &fc, bool &valid) {
    // queue slot i -> pixel of the launch region, in 8x4 tiles (same order as thread_pixel)
    const uint32_t width = fc.width;
    const uint32_t tiles_x = (width + 7u) >> 3;
    const uint32_t w = i >> 5, lane = i & 31u;
    const uint32_t ty = w / tiles_x, tx = w - ty * tiles_x;
    uint32_t x, y;
    valid = region_pixel(fc, tx, ty, lane, x, y);
    return y * width + x;
}

__global__ void __launch_bounds__(256) k_wf_generate(const FrameConsts fc, WavefrontState ws, uint32_t slot_begin, uint32_t n_slots,
                                                     unsigned long long *stats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    uint32_t p = 0;
    if (i < n_slots) p = swizzled_pixel(slot_begin + i, fc, valid);
    if (valid) {
        const uint32_t x = p % fc.width, y = p / fc.width;
        uint32_t rng = tea(p, fc.frame);
        const float jx = next_rand(rng), jy = next_rand(rng);
        const float3 d = primary_dir(fc, (float)x + jx, (float)y + jy);
        ws.ray_o[p] = make_float4(fc.origin.x, fc.origin.y, fc.origin.z, 0.0f);
        ws.ray_d[p] = make_float4(d.x, d.y, d.z, 0.0f);
        ws.thr[p] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u));  // w: depth | sample << 16
        ws.pix[p] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(rng));
    }
    // compact valid pixels into queue 0 (image edges may leave holes in the tile order)
    const uint32_t m = __ballot_sync(0xffffffffu, valid);
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(&ws.counters[0], (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (valid) ws.queue[0][base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = p;
    warp_add_stat(stats, ST_PATHS, valid ? 1u : 0u);
}

// Persistent, warp-cooperative traversal.  Every iteration the warp votes between a NODE step (lanes with
// an internal-child hit intersect one 8-wide node) and a TRIANGLE step (lanes with pending triangle hits
// test one triangle), so the two phases never serialise inside an iteration; triangle groups a lane cannot
// serve yet are postponed on its stack (after Ylitie et al. 2017).  Lanes whose ray has terminated are
// refilled from the queue as soon as WF_FETCH_IDLE lanes are idle (dynamic fetch, Aila & Laine 2009); the
// warp takes rays from the global queue in batches of up to WF_BATCH to keep the single atomic counter cold; short queues
// (late waves, small images, one rank's share of a tile-split frame) use smaller batches so every resident warp gets rays:
// with a fixed 128 a 86 K-ray wave kept 675 of 2 960 warps busy, four rays deep (tools/tile_time.py).
constexpr uint32_t WF_BATCH = 128;

// Resident CTAs per SM the trace kernel is compiled for.  8 (= 64 registers, 12 bytes of spill) for the flattened kernel: with the
// shorter node step of session 3 it beats the 72-register build (3 507 -> 3 574 Mrays/s, profiles/r01_sweep_regs64.txt; the same
// experiment on the session-2 kernel had lost 2 %).  The two-level and cooperative variants would spill ~90 bytes at 64: they keep 7.
#ifndef SOLB_WF_MIN_CTAS
#define SOLB_WF_MIN_CTAS 8
#endif
// TL: two-level scenes.  A lane is either in the TLAS (its "triangle" groups are instance leaves: the triangle step
// enters the instance) or inside a BLAS; the sentinel popped at the end of a BLAS walk reloads the world-space ray.
// COOP (experimental, SOLB_COOP_TRI=1, flattened scenes): instead of a triangle step in which every lane tests ONE of its own
// pending triangles, the warp pools ALL pending triangle tests of its rays and deals them out to all 32 lanes (a lane fetches
// the owner's ray by shuffle), then every owner takes the nearest of its results.  One such step empties every lane's
// triangle group, so the rays that were waiting for a triangle step return to the node phase together: in a replay of the
// headline workload's rays (tools/warp_policy_sim) node steps ran 25.5 instead of 20.6 lanes wide.  The dealing (prefix
// scan, work list, 7 shuffles, ray frame, result gather) costs ~170 instructions per round though, and the replay only wins
// below ~40: measured 3 051 vs 3 300 Mrays/s, so it stays off (profiles/r01_sweep_coop_tri.txt).
template <bool STATS, bool TL, bool COOP>
__global__ void __launch_bounds__(TRACE_BLOCK, (TL || COOP) ? 7 : SOLB_WF_MIN_CTAS) k_wf_trace(const FrameConsts fc, const uint4 *__restrict__ nodes,
                                                          const float4 *__restrict__ tris, const float4 *__restrict__ inst_leaves,
                                                          WavefrontState ws, int qi, unsigned long long *stats, const TraceTuning tune) {
    SOLB_DECL_STACK();
    __shared__ uint32_t s_coop_item[COOP ? TRACE_BLOCK : 1];  // per warp: 32 (owner lane << 27 | triangle index)
    __shared__ float s_coop_t[COOP ? TRACE_BLOCK : 1];        // per warp: hit distance of each dealt test (inf: miss)
    const uint32_t n = ws.counters[qi];
    const uint32_t *__restrict__ queue = qi ? ws.queue[1] : ws.queue[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.counters[qi ^ 1] = 0;  // next wave's output queue
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t nr = 0;
    TraceCounters ctr = { 0, 0 };
    // warp-uniform pool of queue slots [pool_next, pool_end)
    uint32_t pool_next = 0, pool_end = 0;
    bool exhausted = false;
    // per-lane ray state
    bool has_ray = false;
    uint32_t pixel = 0;
    TravRay tr = make_trav_ray(f3(0, 0, 0), f3(0, 0, 1), 0.0f);
    float tmax = 0.0f;
    Hit hit;
    hit.inst = SOLB_MISS; hit.prim = SOLB_MISS; hit.gtri = SOLB_MISS; hit.t = 0.0f; hit.u = 0.0f; hit.v = 0.0f;
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
    bool in_blas = false;          // TL only
    uint32_t cur_inst = SOLB_MISS; // TL only: instance being walked
    uint32_t b_ray = 0;  // lanes holding a ray (warp-uniform)
    for (;;) {
        // ---- refill idle lanes ----
        const uint32_t idle = ~b_ray;
        if (!exhausted && (__popc(idle) >= tune.fetch_idle)) {
            uint32_t want = (uint32_t)__popc(idle);
            uint32_t served = 0;  // idle lanes (in rank order) already given a slot
            while (served < want && !exhausted) {
                if (pool_next >= pool_end) {
                    // recomputed here rather than kept live across the traversal loop
                    const uint32_t batch = min(WF_BATCH, max((uint32_t)tune.min_batch, (n / (gridDim.x * (TRACE_BLOCK / 32)) + 31u) & ~31u));
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&ws.counters[2], batch);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base >= n) { exhausted = true; break; }
                    pool_next = base;
                    pool_end = min(base + batch, n);
                }
                const uint32_t take = min(want - served, pool_end - pool_next);
                const uint32_t rank = (uint32_t)__popc(idle & lt_mask);
                if (!has_ray && rank >= served && rank < served + take) {
                    pixel = queue[pool_next + (rank - served)];
                    const float4 o = ws.ray_o[pixel], d = ws.ray_d[pixel];
                    tr = make_trav_ray(f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), fc.tmin);
                    tmax = fc.tmax;
                    hit.inst = SOLB_MISS; hit.gtri = SOLB_MISS; hit.u = 0.0f; hit.v = 0.0f;
                    ngroup = SOLB_ROOT_GROUP;
                    tgroup = make_uint2(0u, 0u);
                    stack.sp = 0;
                    in_blas = false;
                    has_ray = true;
                    nr++;
                }
                pool_next += take;
                served += take;
            }
            b_ray = __ballot_sync(0xffffffffu, has_ray);
        }
        if (b_ray == 0u) break;  // nothing in flight and nothing left to fetch
        // ---- vote: node step or triangle step ----
        const bool w_node = has_ray && (ngroup.y & 0xff000000u);
        const bool w_tri = has_ray && tgroup.y;
        const uint32_t b_node = __ballot_sync(0xffffffffu, w_node), b_tri = __ballot_sync(0xffffffffu, w_tri);
        const int nn = __popc(b_node), nt = __popc(b_tri);
        if (COOP && !TL) {
            if (nt > 0 && (nn == 0 || __popc(b_tri & ~b_node) >= tune.coop_block)) {
                // ---- cooperative triangle step ----
                uint32_t *work = s_coop_item + (threadIdx.x & ~31u);
                float *res_t = s_coop_t + (threadIdx.x & ~31u);
                const uint32_t mine = w_tri ? tgroup.y : 0u;
                const int cnt = __popc(mine);
                int incl = cnt;
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, dlt);
                    if ((int)lane >= dlt) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31), off = incl - cnt;
                for (int base = 0; base < total; base += 32) {
                    {   // owners publish the tests that fall into this round, nearest-first order of the serial walk
                        uint32_t mm = mine;
                        int k = off - base;
                        while (mm) {
                            const int b = 31 - __clz(mm);
                            mm &= ~(1u << b);
                            if (k >= 0 && k < 32) work[k] = (lane << 27) | (tgroup.x + (uint32_t)b);
                            k++;
                        }
                    }
                    __syncwarp();
                    const int n_work = min(32, total - base);
                    const bool worker = (int)lane < n_work;
                    const uint32_t item = worker ? work[lane] : 0u;
                    const int owner = (int)(item >> 27);
                    const float3 o = f3(__shfl_sync(0xffffffffu, tr.o.x, owner), __shfl_sync(0xffffffffu, tr.o.y, owner),
                                        __shfl_sync(0xffffffffu, tr.o.z, owner));
                    const float3 d = f3(__shfl_sync(0xffffffffu, tr.d.x, owner), __shfl_sync(0xffffffffu, tr.d.y, owner),
                                        __shfl_sync(0xffffffffu, tr.d.z, owner));
                    const float o_tmax = __shfl_sync(0xffffffffu, tmax, owner);
                    float t = 3.4e38f, u = 0.0f, v = 0.0f;
                    uint32_t h_inst = SOLB_MISS, h_gtri = SOLB_MISS;
                    if (worker) {
                        const float4 *tp = tris + (size_t)(item & 0x07ffffffu) * 3;
                        const float4 v0 = SOLB_LDG4(tp + 0), v1 = SOLB_LDG4(tp + 1), v2 = SOLB_LDG4(tp + 2);
                        const RayFrame fr = make_ray_frame(d);
                        float tt, uu, vv;
                        if (intersect_tri(o, d, fr, xyz(v0), xyz(v1), xyz(v2), fc.tmin, o_tmax, tt, uu, vv)) {
                            t = tt; u = uu; v = vv;
                            h_inst = __float_as_uint(v0.w); h_gtri = __float_as_uint(v2.w);
                        }
                        if (STATS) ctr.tris++;
                    }
                    res_t[lane] = t;
                    __syncwarp();
                    // owners: the nearest of their tests in this round (first wins ties, like the serial walk)
                    int best = -1;
                    float best_t = tmax;
                    {
                        const int lo = max(off - base, 0), hi = min(off - base + cnt, 32);
                        for (int k = lo; k < hi; k++) {
                            const float tk = res_t[k];
                            if (tk < best_t) { best_t = tk; best = k; }
                        }
                    }
                    const int src = best >= 0 ? best : (int)lane;
                    const float bu = __shfl_sync(0xffffffffu, u, src), bv = __shfl_sync(0xffffffffu, v, src);
                    const uint32_t bi = __shfl_sync(0xffffffffu, h_inst, src), bg = __shfl_sync(0xffffffffu, h_gtri, src);
                    if (best >= 0) {
                        tmax = best_t;
                        hit.t = best_t; hit.u = bu; hit.v = bv; hit.inst = bi; hit.gtri = bg;
                    }
                    __syncwarp();
                }
                if (w_tri) tgroup.y = 0u;
            } else if (w_node) {
                if (tgroup.y) stack.push(tgroup);  // postpone the pending triangles
                trav_node_step(nodes, tr, tmax, ngroup, tgroup, stack);
                if (STATS) ctr.nodes++;
            }
        } else if (nt > 0 && nt * tune.tri_weight >= nn * tune.node_weight) {
            if (w_tri) {
                if (TL && !in_blas) {
                    trav_enter_instance(inst_leaves, tr.o, tr.d, tr, ngroup, tgroup, cur_inst, stack);
                    in_blas = true;
                } else {
                    if (trav_tri_step(tris, tr, tmax, tgroup, hit) && TL) hit.inst = cur_inst;
                    if (STATS) ctr.tris++;
                }
            }
        } else if (w_node) {
            if (tgroup.y) stack.push(tgroup);  // postpone the pending triangles
            trav_node_step(nodes, tr, tmax, ngroup, tgroup, stack);
            if (STATS) ctr.nodes++;
        }
        // ---- lanes with nothing in hand: pop, or finish the ray ----
        if (has_ray && !(ngroup.y & 0xff000000u) && !tgroup.y) {
            if (stack.empty()) {
                ws.hit[pixel] = make_uint4(hit.inst, hit.gtri, __float_as_uint(hit.u), __float_as_uint(hit.v));
                has_ray = false;
            } else {
                const uint2 e = stack.pop();
                if (TL && e.y == 0u) {  // sentinel: back to the TLAS with the world-space ray
                    const float4 o = ws.ray_o[pixel], d = ws.ray_d[pixel];
                    tr = make_trav_ray(f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), fc.tmin);
                    in_blas = false;
                } else if (e.y & 0xff000000u) ngroup = e;
                else tgroup = e;
            }
        }
        b_ray = __ballot_sync(0xffffffffu, has_ray);
    }
    warp_add_stat(stats, ST_RAYS, nr);
    if (STATS) {
        warp_add_stat(stats, ST_NODES, ctr.nodes);
        warp_add_stat(stats, ST_TRIS, ctr.tris);
    }
}

// ---------------------------------------------------------------------------------------------------
// Ray-pool traversal kernel.  k_wf_trace binds one ray to one lane, so a lane idles whenever the warp's vote
// picks the step type its ray does not need (measured: 20 of 32 lanes in node steps, 16.5 in triangle steps,
// profiles/r01_ncu_k_wf_trace_c.txt).  Here every warp owns PL_SLOTS rays whose traversal state lives in
// shared memory (SoA over slots); three warp-uniform 64-bit masks say which slots are free, need a node step,
// or need a triangle step.  Each iteration the warp picks the fuller class and hands its slots to lanes
// (lane i takes the i-th set bit), so node / triangle / ray-setup steps all run with (nearly) full warps.
constexpr int PL_SLOTS = 64;   // rays per warp
constexpr int PL_STACK = 4;    // stack entries per slot kept in shared memory
constexpr int PL_SPILL = 60;   // further entries per slot in global memory
constexpr int PL_WARPS = TRACE_BLOCK / 32;
enum { PL_NODE = 1, PL_TRI = 2, PL_FREE = 4 };

struct PoolWarp {
    float ox[PL_SLOTS], oy[PL_SLOTS], oz[PL_SLOTS], dx[PL_SLOTS], dy[PL_SLOTS], dz[PL_SLOTS];
    float tmax[PL_SLOTS];
    uint32_t pixel[PL_SLOTS], hit_inst[PL_SLOTS], hit_gtri[PL_SLOTS];
    float hit_u[PL_SLOTS], hit_v[PL_SLOTS];
    uint32_t ngx[PL_SLOTS], ngy[PL_SLOTS], tgx[PL_SLOTS], tgy[PL_SLOTS];
    uint2 stack[PL_STACK][PL_SLOTS];
    uint8_t sp[PL_SLOTS];
    uint8_t status[PL_SLOTS];  // PL_NODE | PL_TRI, or PL_FREE
    uint8_t assign[32];        // slot handled by lane i in the current step
};

struct PoolStack {
    PoolWarp *P;
    uint2 *spill;  // this slot's global spill area
    int slot, sp;
    __device__ __forceinline__ void push(uint2 v) {
        if (sp < PL_STACK) P->stack[sp][slot] = v;
        else if (sp < PL_STACK + PL_SPILL) spill[sp - PL_STACK] = v;
        sp++;
    }
    __device__ __forceinline__ uint2 pop() {
        sp--;
        return sp < PL_STACK ? P->stack[sp][slot] : spill[sp - PL_STACK];
    }
    __device__ __forceinline__ bool empty() const { return sp == 0; }
};

// Hand the first `cnt` slots of class mask (lo | hi << 32) to lanes 0..cnt-1: every lane looks at its own two slots
// (lane, lane + 32), computes their rank inside the class with a prefix popcount and scatters them through
// shared memory.  Returns this lane's slot or -1.
__device__ __forceinline__ int pool_assign(PoolWarp &P, uint32_t lo, uint32_t hi, int cnt, int lane, uint32_t lt_mask) {
    const int r0 = __popc(lo & lt_mask), r1 = __popc(lo) + __popc(hi & lt_mask);
    if (((lo >> lane) & 1u) && r0 < cnt) P.assign[r0] = (uint8_t)lane;
    if (((hi >> lane) & 1u) && r1 < cnt) P.assign[r1] = (uint8_t)(lane + 32);
    __syncwarp();
    const int slot = lane < cnt ? (int)P.assign[lane] : -1;
    __syncwarp();
    return slot;
}

template <bool STATS>
__global__ void __launch_bounds__(TRACE_BLOCK) k_wf_trace_pool(const FrameConsts fc, const uint4 *__restrict__ nodes,
                                                               const float4 *__restrict__ tris, WavefrontState ws, int qi,
                                                               unsigned long long *stats, const TraceTuning tune, uint2 *spill_base) {
    __shared__ PoolWarp s_pool[PL_WARPS];
    PoolWarp &P = s_pool[threadIdx.x >> 5];
    const uint32_t n = ws.counters[qi];
    const uint32_t *__restrict__ queue = qi ? ws.queue[1] : ws.queue[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.counters[qi ^ 1] = 0;  // next wave's output queue
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const size_t gwarp = (size_t)blockIdx.x * PL_WARPS + (threadIdx.x >> 5);
    uint2 *warp_spill = spill_base + gwarp * (size_t)(PL_SLOTS * PL_SPILL);
    uint32_t nr = 0;
    TraceCounters ctr = { 0, 0 };
    P.status[lane] = PL_FREE;
    P.status[lane + 32] = PL_FREE;
    __syncwarp();
    bool exhausted = false;
    for (;;) {
        // ---- slot classes from the per-slot status bytes (each lane reports its two slots) ----
        const uint32_t st0 = P.status[lane], st1 = P.status[lane + 32];
        const uint32_t node_lo = __ballot_sync(0xffffffffu, st0 & PL_NODE), node_hi = __ballot_sync(0xffffffffu, st1 & PL_NODE);
        const uint32_t tri_lo = __ballot_sync(0xffffffffu, st0 & PL_TRI), tri_hi = __ballot_sync(0xffffffffu, st1 & PL_TRI);
        const uint32_t free_lo = __ballot_sync(0xffffffffu, st0 & PL_FREE), free_hi = __ballot_sync(0xffffffffu, st1 & PL_FREE);
        const int nn = __popc(node_lo) + __popc(node_hi), nt = __popc(tri_lo) + __popc(tri_hi);
        const int nfree = __popc(free_lo) + __popc(free_hi);
        // ---- ray setup step: fill free slots from the queue ----
        if (!exhausted && (nfree >= tune.pool_refill || nn + nt == 0)) {
            const int want = nfree < 32 ? nfree : 32;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ws.counters[2], (uint32_t)want);
            base = __shfl_sync(0xffffffffu, base, 0);
            const int got = base >= n ? 0 : (int)min((uint32_t)want, n - base);
            if (got < want) exhausted = true;
            const int slot = pool_assign(P, free_lo, free_hi, got, lane, lt_mask);
            if (slot >= 0) {
                const uint32_t pixel = queue[base + lane];
                const float4 o = ws.ray_o[pixel], d = ws.ray_d[pixel];
                P.ox[slot] = o.x; P.oy[slot] = o.y; P.oz[slot] = o.z;
                P.dx[slot] = d.x; P.dy[slot] = d.y; P.dz[slot] = d.z;
                P.tmax[slot] = fc.tmax;
                P.pixel[slot] = pixel;
                P.hit_inst[slot] = SOLB_MISS; P.hit_gtri[slot] = SOLB_MISS; P.hit_u[slot] = 0.0f; P.hit_v[slot] = 0.0f;
                P.ngx[slot] = 0u; P.ngy[slot] = 0x80000000u;  // SOLB_ROOT_GROUP
                P.tgx[slot] = 0u; P.tgy[slot] = 0u;
                P.sp[slot] = 0;
                P.status[slot] = PL_NODE;
                nr++;
            }
            __syncwarp();
            if (got > 0) continue;  // re-read the classes
        }
        if (nn + nt == 0) break;  // nothing in flight, nothing left to fetch
        // ---- pick the fuller class and hand its slots to lanes ----
        const bool do_tri = nt > 0 && (nt * tune.tri_weight >= nn * tune.node_weight || nt >= 32);
        const int cnt = do_tri ? (nt < 32 ? nt : 32) : (nn < 32 ? nn : 32);
        const int slot = pool_assign(P, do_tri ? tri_lo : node_lo, do_tri ? tri_hi : node_hi, cnt, lane, lt_mask);
        if (slot >= 0) {
            PoolStack stack;
            stack.P = &P;
            stack.slot = slot;
            stack.spill = warp_spill + (size_t)slot * PL_SPILL;
            stack.sp = P.sp[slot];
            uint2 ngroup = make_uint2(P.ngx[slot], P.ngy[slot]);
            uint2 tgroup = make_uint2(P.tgx[slot], P.tgy[slot]);
            TravRay tr;
            tr.o = f3(P.ox[slot], P.oy[slot], P.oz[slot]);
            tr.d = f3(P.dx[slot], P.dy[slot], P.dz[slot]);
            tr.tmin = fc.tmin;
            if (do_tri) {
                tr.frame = make_ray_frame(tr.d);
                float tmax = P.tmax[slot];
                Hit hit;
                hit.inst = SOLB_MISS;
                hit.gtri = P.hit_gtri[slot];
                trav_tri_step(tris, tr, tmax, tgroup, hit);
                if (STATS) ctr.tris++;
                if (hit.inst != SOLB_MISS) {
                    P.tmax[slot] = tmax;
                    P.hit_inst[slot] = hit.inst; P.hit_gtri[slot] = hit.gtri; P.hit_u[slot] = hit.u; P.hit_v[slot] = hit.v;
                }
            } else {
                tr.idir = f3(safe_rcp_dir(tr.d.x), safe_rcp_dir(tr.d.y), safe_rcp_dir(tr.d.z));
                set_trav_octant(tr, tr.d);
                if (tgroup.y) stack.push(tgroup);  // postpone the pending triangles
                trav_node_step(nodes, tr, P.tmax[slot], ngroup, tgroup, stack);
                if (STATS) ctr.nodes++;
            }
            // nothing in hand: pop, or finish the ray
            uint32_t status = 0;
            if (!(ngroup.y & 0xff000000u) && !tgroup.y) {
                if (stack.empty()) {
                    ws.hit[P.pixel[slot]] = make_uint4(P.hit_inst[slot], P.hit_gtri[slot], __float_as_uint(P.hit_u[slot]),
                                                       __float_as_uint(P.hit_v[slot]));
                    status = PL_FREE;
                } else {
                    const uint2 e = stack.pop();
                    if (e.y & 0xff000000u) ngroup = e; else tgroup = e;
                }
            }
            if (ngroup.y & 0xff000000u) status |= PL_NODE;
            if (tgroup.y) status |= PL_TRI;
            P.ngx[slot] = ngroup.x; P.ngy[slot] = ngroup.y;
            P.tgx[slot] = tgroup.x; P.tgy[slot] = tgroup.y;
            P.sp[slot] = (uint8_t)stack.sp;
            P.status[slot] = (uint8_t)status;
        }
        __syncwarp();
    }
    warp_add_stat(stats, ST_RAYS, nr);
    if (STATS) {
        warp_add_stat(stats, ST_NODES, ctr.nodes);
        warp_add_stat(stats, ST_TRIS, ctr.tris);
    }
}

// SORT: material-sorted shading.  Each block takes 256 queue entries, classifies their hit records (miss /
// emissive / surface material bucket) and counting-sorts the entries by class in shared memory before shading, so
// the lanes of a warp run the same rmiss / emissive / BRDF path of pathtrace.rchit.
template <bool SORT>
__global__ void __launch_bounds__(256) k_wf_shade(const FrameConsts fc, const DeviceInstance *__restrict__ instances,
                                                  const ShadeRecord *__restrict__ shade, WavefrontState ws, int qi,
                                                  unsigned long long *stats) {
    __shared__ uint32_t s_count[8], s_perm[256];
    const uint32_t n = ws.counters[qi];
    const uint32_t *__restrict__ queue_in = ws.queue[qi];
    uint32_t *__restrict__ queue_out = ws.queue[qi ^ 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.counters[2] = 0;  // trace work head of the next wave
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t nh = 0, np = 0;
    // grid-stride over whole 256-entry chunks so the block-wide sort and the warp ballot always see whole blocks
    const uint32_t n_round = (n + 255u) & ~255u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool alive = false;
        uint32_t p = 0;
        bool have = i < n;
        if (have) p = queue_in[i];
        if (SORT) {
            if (threadIdx.x < 8) s_count[threadIdx.x] = 0;
            __syncthreads();
            uint32_t cls = 7, rank = 0;
            if (have) {
                const uint32_t inst = ws.hit[p].x;
                if (inst == SOLB_MISS) cls = 0;
                else {
                    const DeviceInstance &in = instances[inst];
                    cls = (in.mat[4] >= 1.0f || in.mat[5] >= 1.0f || in.mat[6] >= 1.0f) ? 1u : 2u + in.material % 5u;
                }
                rank = atomicAdd(&s_count[cls], 1u);
            }
            __syncthreads();
            if (have) {
                uint32_t off = 0;
                for (uint32_t c = 0; c < cls; c++) off += s_count[c];
                s_perm[off + rank] = p;
            }
            __syncthreads();
            uint32_t total = 0;
            for (int c = 0; c < 7; c++) total += s_count[c];
            have = threadIdx.x < total;
            if (have) p = s_perm[threadIdx.x];
            __syncthreads();
        }
        if (have) {
            const uint4 h = ws.hit[p];
            const float4 t4 = ws.thr[p], x4 = ws.pix[p];
            float3 thr = f3(t4.x, t4.y, t4.z), pixel = f3(x4.x, x4.y, x4.z);
            uint32_t ds = __float_as_uint(t4.w), rng = __float_as_uint(x4.w);
            uint32_t depth = ds & 0xffffu, sample = ds >> 16;
            const float4 d4 = ws.ray_d[p];
            float3 o = f3(0, 0, 0), d = f3(d4.x, d4.y, d4.z);
            bool end_path;
            if (h.x != SOLB_MISS) {
                nh++;
                float3 hv;
                const bool done = shade_hit(instances, shade, fc.texb, h.x, h.y, __uint_as_float(h.z), __uint_as_float(h.w), o, d, rng, hv);
                depth++;
                thr = thr * hv;
                end_path = done;
                if (!done && depth > fc.max_bounces) { thr = f3(0, 0, 0); end_path = true; }
            } else {
                thr = thr * shade_miss(fc.enable_sky, d);
                end_path = true;
            }
            alive = true;
            if (end_path) {
                pixel = pixel + thr;
                sample++;
                if (sample < fc.spp) {
                    const uint32_t x = p % fc.width, y = p / fc.width;
                    const float jx = next_rand(rng), jy = next_rand(rng);
                    o = fc.origin;
                    d = primary_dir(fc, (float)x + jx, (float)y + jy);
                    thr = f3(1, 1, 1);
                    depth = 0;
                    np++;
                } else {
                    alive = false;
                }
                ws.pix[p] = make_float4(pixel.x, pixel.y, pixel.z, __uint_as_float(rng));
            } else {
                ws.pix[p].w = __uint_as_float(rng);
            }
            if (alive) {
                ws.ray_o[p] = make_float4(o.x, o.y, o.z, 0.0f);
                ws.ray_d[p] = make_float4(d.x, d.y, d.z, 0.0f);
                ws.thr[p] = make_float4(thr.x, thr.y, thr.z, __uint_as_float(depth | (sample << 16)));
            }
        }
        // warp-ballot compaction + one atomic per warp to append survivors to the next queue
        const uint32_t m = __ballot_sync(0xffffffffu, alive);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(&ws.counters[qi ^ 1], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (alive) queue_out[base + __popc(m & ((1u << lane) - 1u))] = p;
    }
    warp_add_stat(stats, ST_HITS, nh);
    warp_add_stat(stats, ST_PATHS, np);
}

__global__ void __launch_bounds__(256) k_wf_resolve(const FrameConsts fc, WavefrontState ws, float4 *accum, uint32_t *render,
                                                    uint32_t slot_begin, uint32_t n_slots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    bool valid;
    const uint32_t p = swizzled_pixel(slot_begin + i, fc, valid);
    if (!valid) return;
    const float4 x4 = ws.pix[p];
    uint32_t rgba;
    const float4 out = resolve_pixel(fc, f3(x4.x, x4.y, x4.z), accum[p], rgba);
    accum[p] = out;
    if (render) render[p] = rgba;
}

// ---------------------------------------------------------------------------------------------------
// 5-pathtrace, warp-local wavefront schedule (SOLB_SCHEDULE_WARPFRONT): ONE persistent kernel per frame.
//
// The queue-based wavefront above pays for its global queues three times: every one of the ~72 waves of a frame ends in a drain
// tail (a persistent warp holds ~7 rays per lane and wave, so the last ray of each lane runs in a thinning warp), the shade kernel
// is a separate memory-bound launch between two traversal launches, and the host has to poll survivor counts to know when to stop.
// Here every WARP is its own wavefront machine over a pool of WL_POOL pixels it owns:
//   * its 32 lanes walk one ray each with the voted node / triangle steps of k_wf_trace;
//   * a finished ray's pixel slot goes on the warp's to-shade list (shared memory, ballot-compacted);
//   * once 32 slots wait there, the whole warp runs ONE shade step over them - pathtrace.rchit / rmiss / bounce cap / next
//     sample exactly as k_wf_shade does, 32 lanes wide and converged - and the surviving slots go on the warp's ready list;
//   * idle lanes take ready slots; a pixel that has finished its samples is resolved into the accumulation / render targets in
//     the shade step and its slot is refilled with the next pixel of the frame's cursor (8x4 tiles, one atomic per tile).
// No global queue, no inter-warp communication, no waves: the only drain is the end of the frame.  Path state lives in
// slot-indexed arrays sized by the grid (36 MB), so it stays in L2.  Per-pixel arithmetic, RNG order and accumulation are those of
// k_wf_generate / k_wf_shade / k_wf_resolve: the frames are bit-identical to the other schedules'.
// Path-state accesses bypass L1 (ld.global.cg / st.global.cg): the records stream through once per ray, while the 128-256 KB of
// L1 are what keeps node and triangle fetches at an 88 % hit rate.  SOLB_WL_CG=0 restores default caching for A/B runs.
#ifndef SOLB_WL_CG
#define SOLB_WL_CG 1
#endif
template <class T>
__device__ __forceinline__ T wl_ld(const T *p) {
#if SOLB_WL_CG
    return __ldcg(p);
#else
    return *p;
#endif
}
template <class T>
__device__ __forceinline__ void wl_st(T *p, const T &v) {
#if SOLB_WL_CG
    __stcg(p, v);
#else
    *p = v;
#endif
}

// a new ray of slot gs: origin (+ pixel id), direction and the traversal set-up derived from the direction
__device__ __forceinline__ void wl_store_ray(const WarpfrontState &wl, uint32_t gs, float3 o, float3 d, uint32_t pixel_id) {
    const float3 idir = f3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    const RayFrame fr = make_ray_frame(d);
    wl_st(wl.ray_o + gs, make_float4(o.x, o.y, o.z, __uint_as_float(pixel_id)));
    wl_st(wl.ray_d + gs, make_float4(d.x, d.y, d.z, 0.0f));
    wl_st(wl.ray_i + gs, make_float4(idir.x, idir.y, idir.z, 0.0f));
    wl_st(wl.frame0 + gs, make_float4(fr.e1.x, fr.e1.y, fr.e1.z, fr.e2.x));
    wl_st(wl.frame1 + gs, make_float4(fr.e2.y, fr.e2.z, 0.0f, 0.0f));
}

__device__ __forceinline__ void wl_push(uint8_t *list, uint32_t &n, bool flag, uint32_t slot, uint32_t lt_mask) {
    const uint32_t m = __ballot_sync(0xffffffffu, flag);
    if (flag) list[n + __popc(m & lt_mask)] = (uint8_t)slot;
    n += (uint32_t)__popc(m);
}

// per-warp control block in shared memory: the three slot lists + the claimed range of the frame's cursor
// Both work lists are FIFO.  With LIFO lists (the first version) a slot at the bottom of the ready stack was only served once
// everything above it had finished, so the pixels of a pool drifted apart and, once the frame's cursor ran dry, the stale ones
// were finished one after the other by a handful of lanes: the drain of a warp took several pixel lifetimes.  Served in
// arrival order all pixels of a pool advance at the same rate and the drain is one pixel lifetime.
constexpr uint32_t WL_RING = 128;  // ready ring: power of two >= WL_POOL
static_assert(WL_POOL <= (int)WL_RING && WL_POOL % 32 == 0, "WL_POOL must be a multiple of 32 and fit the ready ring");
struct WlWarp {
    uint8_t ready[WL_RING];  // ring: entries head .. head + n_ready - 1 (mod WL_RING)
    uint8_t shade[WL_POOL];  // queue: entries 0 .. n_shade - 1, oldest first
    uint8_t free_[WL_POOL];  // stack
    uint32_t pool_next, pool_end;
};
// list state travels in ONE word: n_ready | n_shade << 8 | n_free << 16 | exhausted << 24 | ready head << 25
__device__ __forceinline__ uint32_t wl_pack(uint32_t n_ready, uint32_t n_shade, uint32_t n_free, bool exhausted, uint32_t head) {
    return n_ready | (n_shade << 8) | (n_free << 16) | (exhausted ? 1u << 24 : 0u) | (head << 25);
}
__device__ __forceinline__ void wl_push_ready(WlWarp *W, uint32_t head, uint32_t &n, bool flag, uint32_t slot, uint32_t lt_mask) {
    const uint32_t m = __ballot_sync(0xffffffffu, flag);
    if (flag) W->ready[(head + n + (uint32_t)__popc(m & lt_mask)) & (WL_RING - 1u)] = (uint8_t)slot;
    n += (uint32_t)__popc(m);
}

// Generate step: up to 32 free slots take the next pixels of the frame (k_wf_generate).
__device__ __forceinline__ uint32_t wl_generate_step(const FrameConsts *fcp, const WarpfrontState *wlp, WlWarp *W, uint32_t slot_base,
                                                  uint32_t counts, uint32_t n_region_slots, uint32_t batch, uint32_t *ctr) {
    const FrameConsts &fc = *fcp;
    const WarpfrontState &wl = *wlp;
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    uint32_t n_ready = counts & 0xffu, n_shade = (counts >> 8) & 0xffu, n_free = (counts >> 16) & 0xffu;
    bool exhausted = (counts >> 24) & 1u;
    const uint32_t head = counts >> 25;
    const uint32_t cnt = min(n_free, 32u);
    const bool mine = lane < cnt;
    const uint32_t my_slot = mine ? (uint32_t)W->free_[n_free - 1u - lane] : 0u;
    n_free -= cnt;
    uint32_t pool_next = W->pool_next, pool_end = W->pool_end;
    __syncwarp();
    // claim cnt frame slots (warp-uniform bookkeeping, like the dynamic fetch of k_wf_trace)
    uint32_t my_idx = 0xffffffffu, served = 0;
    while (served < cnt && !exhausted) {
        if (pool_next >= pool_end) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(wl.cursor, batch);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n_region_slots) { exhausted = true; break; }
            pool_next = base;
            pool_end = min(base + batch, n_region_slots);
        }
        const uint32_t take = min(cnt - served, pool_end - pool_next);
        if (mine && lane >= served && lane < served + take) my_idx = pool_next + (lane - served);
        pool_next += take;
        served += take;
    }
    if (lane == 0) { W->pool_next = pool_next; W->pool_end = pool_end; }
    bool valid = false, retry = false;
    if (my_idx != 0xffffffffu) {
        const uint32_t p = swizzled_pixel(my_idx, fc, valid);
        if (valid) {
            const uint32_t x = p % fc.width, y = p / fc.width;
            uint32_t rng = tea(p, fc.frame);                       // pathtrace.rgen:47
            const float jx = next_rand(rng), jy = next_rand(rng);  // :52
            const float3 d = primary_dir(fc, (float)x + jx, (float)y + jy);
            const uint32_t gs = slot_base + my_slot;
            wl_store_ray(wl, gs, fc.origin, d, p);
            wl_st(wl.thr + gs, make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u)));
            wl_st(wl.pix + gs, make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(rng)));
            ctr[1]++;
        } else {
            retry = true;  // a hole of the tile order (image edge): the slot stays free
        }
    }
    wl_push_ready(W, head, n_ready, valid, my_slot, lt_mask);
    wl_push(W->free_, n_free, retry, my_slot, lt_mask);
    if (exhausted) n_free = 0u;  // nothing left to start: free slots are retired
    __syncwarp();
    return wl_pack(n_ready, n_shade, n_free, exhausted, head);
}

// Shade step: up to 32 finished rays, one per lane (k_wf_shade + k_wf_resolve).
__device__ __forceinline__ uint32_t wl_shade_step(const FrameConsts *fcp, const DeviceInstance *__restrict__ instances,
                                               const ShadeRecord *__restrict__ shade, const WarpfrontState *wlp, float4 *frame_sum,
                                               WlWarp *W, uint32_t slot_base, uint32_t counts, uint32_t *ctr) {
    const FrameConsts &fc = *fcp;
    const WarpfrontState &wl = *wlp;
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    uint32_t n_ready = counts & 0xffu, n_shade = (counts >> 8) & 0xffu, n_free = (counts >> 16) & 0xffu;
    const bool exhausted = (counts >> 24) & 1u;
    const uint32_t head = counts >> 25;
    const uint32_t cnt = min(n_shade, 32u);
    const bool mine = lane < cnt;
    const uint32_t my_slot = mine ? (uint32_t)W->shade[lane] : 0u;  // the oldest cnt entries
    n_shade -= cnt;
    const uint32_t moved = lane < n_shade ? (uint32_t)W->shade[cnt + lane] : 0u;  // at most 31 younger ones move to the front
    __syncwarp();
    if (lane < n_shade) W->shade[lane] = (uint8_t)moved;
    bool alive = false, freed = false;
    if (mine) {
        const uint32_t gs = slot_base + my_slot;
        const uint4 h = wl_ld(wl.hit + gs);
        const float4 t4 = wl_ld(wl.thr + gs), x4 = wl_ld(wl.pix + gs), d4 = wl_ld(wl.ray_d + gs);
        const uint32_t p = __float_as_uint(wl_ld(&wl.ray_o[gs].w));
        float3 thr = f3(t4.x, t4.y, t4.z), pixel = f3(x4.x, x4.y, x4.z);
        const uint32_t ds = __float_as_uint(t4.w);
        uint32_t rng = __float_as_uint(x4.w), depth = ds & 0xffffu, sample = ds >> 16;
        float3 o = f3(0, 0, 0), d = f3(d4.x, d4.y, d4.z);
        bool end_path;
        ctr[2]++;  // one shaded record == one traced ray
        if (h.x != SOLB_MISS) {
            ctr[0]++;
            float3 hv;
            const bool done = shade_hit(instances, shade, fc.texb, h.x, h.y, __uint_as_float(h.z), __uint_as_float(h.w), o, d, rng, hv);
            depth++;
            thr = thr * hv;  // pathtrace.rgen:77
            end_path = done;
            if (!done && depth > fc.max_bounces) { thr = f3(0, 0, 0); end_path = true; }  // :81-84
        } else {
            thr = thr * shade_miss(fc.enable_sky, d);
            end_path = true;
        }
        alive = true;
        if (end_path) {
            pixel = pixel + thr;  // :86
            sample++;
            if (sample < fc.spp) {
                const uint32_t x = p % fc.width, y = p / fc.width;
                const float jx = next_rand(rng), jy = next_rand(rng);
                o = fc.origin;
                d = primary_dir(fc, (float)x + jx, (float)y + jy);
                thr = f3(1, 1, 1);
                depth = 0;
                ctr[1]++;
            } else {
                alive = false;
                freed = true;
                // the sum of the pixel's samples; pathtrace.rgen:88-103 (mean, running mix, gamma) runs in k_wf_resolve on the ctx
                // stream, so that this kernel touches neither target and consecutive frames can overlap (launch wrapper)
                wl_st(frame_sum + p, make_float4(pixel.x, pixel.y, pixel.z, __uint_as_float(rng)));
            }
        }
        if (alive) {
            wl_store_ray(wl, gs, o, d, p);
            wl_st(wl.thr + gs, make_float4(thr.x, thr.y, thr.z, __uint_as_float(depth | (sample << 16))));
            wl_st(wl.pix + gs, make_float4(pixel.x, pixel.y, pixel.z, __uint_as_float(rng)));
        }
    }
    wl_push_ready(W, head, n_ready, alive, my_slot, lt_mask);
    if (!exhausted) wl_push(W->free_, n_free, freed, my_slot, lt_mask);
    __syncwarp();
    return wl_pack(n_ready, n_shade, n_free, exhausted, head);
}

// list sizes live in ONE register in the traversal loop too: the kernel sits at the 64-register edge
#define WL_N_READY(c) ((c) & 0xffu)
#define WL_N_SHADE(c) (((c) >> 8) & 0xffu)
#define WL_N_FREE(c) (((c) >> 16) & 0xffu)
#define WL_EXHAUSTED(c) (((c) >> 24) & 1u)
#define WL_NO_SLOT 0xffu

// Register budget.  At 64 registers (8 resident CTAs) the first version kept b_ray / has_ray and the per-ray constants in local
// memory and re-read special registers for addresses: 275 warp-instructions per ray against ~225 for the queue-based pair of
// kernels (profiles/r02_ncu_k_pt_warpfront_a.txt).  Per-lane values that are written once per ray or once per accepted hit
// and read in one place only therefore live in shared memory, [field][thread] so a warp's 16-byte accesses are conflict-free:
//   s_hit   : the hit record (instance, triangle, u, v): written when a triangle test is accepted, read when the ray finishes
//   s_frame : the ray-space frame of the watertight test (6 floats): written at refill, read by the triangle step
template <bool STATS, bool TL>
__global__ void __launch_bounds__(WL_BLOCK, (TL ? 7 : SOLB_WF_MIN_CTAS) * (TRACE_BLOCK / WL_BLOCK))
k_pt_warpfront(const __grid_constant__ FrameConsts fc, const uint4 *__restrict__ nodes, const float4 *__restrict__ tris,
               const float4 *__restrict__ inst_leaves, const DeviceInstance *__restrict__ instances,
               const ShadeRecord *__restrict__ shade, const __grid_constant__ WarpfrontState wl, float4 *frame_sum,
               unsigned long long *stats, const uint32_t n_region_slots, const uint32_t pool_limit,
               const __grid_constant__ TraceTuning tune) {
    SOLB_DECL_STACK_N(WL_BLOCK);
    // opaque to the optimiser: otherwise ptxas re-derives the column address from tid / the shared window base at every push and
    // pop (2 x S2R + MOV + LEA + IMAD) instead of keeping the register (k_wf_trace, with less room, is better off re-deriving)
    asm volatile("" : "+r"(stack.sm));
    __shared__ WlWarp s_warp[WL_BLOCK / 32];
    __shared__ uint4 s_hit[WL_BLOCK];
    __shared__ float4 s_frame0[WL_BLOCK];  // e1.xyz, e2.x
    __shared__ float2 s_frame1[WL_BLOCK];  // e2.yz
    const uint32_t tid = threadIdx.x, lane = tid & 31u, lt_mask = (1u << lane) - 1u;
    WlWarp *const W = &s_warp[tid >> 5];
    const uint32_t slot_base = (blockIdx.x * (WL_BLOCK / 32) + (tid >> 5)) * (uint32_t)WL_POOL;  // < 2^32
    if (slot_base >= wl.n_warps * (uint32_t)WL_POOL) return;
    // pool_limit <= WL_POOL slots are used: small regions (a rank's share of a tile-split frame, small images) are shared out
    // evenly over the warps instead of the first warps claiming WL_POOL pixels each and the rest finding the cursor exhausted
    uint32_t counts = wl_pack(0u, 0u, pool_limit, false, 0u);  // warp-uniform
    for (uint32_t i = lane; i < pool_limit; i += 32u) W->free_[i] = (uint8_t)i;
    if (lane == 0) { W->pool_next = 0u; W->pool_end = 0u; }
    __syncwarp();
    uint32_t step_ctr[3] = { 0u, 0u, 0u };  // hits, paths, rays: counted inside the service steps
    TraceCounters ctr = { 0, 0 };
    // per-lane ray state (as k_wf_trace, minus the hit record and the frame).  A lane holds a ray in flight exactly while it
    // has a node group or a triangle group in hand or entries on its stack (WL_ACTIVE): no separate flag.
#define WL_ACTIVE() (((ngroup.y & 0xff000000u) | tgroup.y | (uint32_t)stack.sp) != 0u)
    uint32_t slot = WL_NO_SLOT;  // pool slot this lane holds: its ray in flight, or finished and not yet collected
    TravRay tr = make_trav_ray(f3(0, 0, 0), f3(0, 0, 1), fc.tmin);
    float tmax = 0.0f;
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
    bool in_blas = false;
    uint32_t cur_inst = SOLB_MISS;
    uint32_t b_ray = 0;
    const uint32_t fetch_idle = (uint32_t)tune.wl_fetch_idle, gen_min = min((uint32_t)tune.wl_gen_min, max(8u, pool_limit - 32u));
    for (;;) {
        const uint32_t n_idle = (uint32_t)__popc(~b_ray);
        // anything to service?  (enough idle lanes | enough free slots; free is 0 once exhausted)
        if (n_idle >= fetch_idle || WL_N_FREE(counts) >= gen_min) {
            // ---- collect: lanes whose ray has finished since the last service hand their slot to the shade list ----
            // (deferred to here: a finished lane just idles, so the per-iteration tail of the loop is a pop or nothing; doing the
            // hit-record store, the ballot and the list append in every iteration cost 45 instructions in 60 % of the iterations)
            {
                const bool done = !WL_ACTIVE() && slot != WL_NO_SLOT;
                const uint32_t b_done = __ballot_sync(0xffffffffu, done);
                if (b_done) {
                    if (done) {
                        wl_st(wl.hit + (slot_base + slot), s_hit[tid]);
                        W->shade[WL_N_SHADE(counts) + __popc(b_done & lt_mask)] = (uint8_t)slot;
                        slot = WL_NO_SLOT;
                    }
                    counts += (uint32_t)__popc(b_done) << 8;
                    __syncwarp();
                }
            }
            const bool idle_enough = n_idle >= fetch_idle;
            // a partial shade / generate step (fewer than 32 slots) only when the warp is really short of rays: these steps cost
            // ~600 instructions whatever their width
            const bool starving = n_idle >= (uint32_t)tune.wl_starve_idle;
            // ---- service steps (once per ~32 rays): generate + shade ----
            // The traversal state is parked in (volatile) local memory around them, so that no per-ray value is live across the
            // shading code and the register allocator can give the traversal loop the whole file.
            const bool want_gen = WL_N_FREE(counts) > 0u &&
                (WL_N_FREE(counts) >= gen_min || (starving && WL_N_READY(counts) == 0u && WL_N_SHADE(counts) < 32u));
            const bool want_shade = WL_N_SHADE(counts) >= 32u || (WL_N_SHADE(counts) > 0u && WL_N_READY(counts) == 0u && starving);
            if (want_gen || want_shade) {
                volatile uint32_t park[24];
                park[0] = f2u(tr.o.x); park[1] = f2u(tr.o.y); park[2] = f2u(tr.o.z);
                park[3] = f2u(tr.d.x); park[4] = f2u(tr.d.y); park[5] = f2u(tr.d.z);
                park[6] = f2u(tr.idir.x); park[7] = f2u(tr.idir.y); park[8] = f2u(tr.idir.z);
                park[9] = tr.oct_inv; park[10] = tr.pow4_lo; park[11] = tr.pow4_hi;
                park[12] = f2u(tmax); park[13] = ngroup.x; park[14] = ngroup.y; park[15] = tgroup.x; park[16] = tgroup.y;
                park[17] = slot; park[18] = (uint32_t)stack.sp; park[19] = cur_inst; park[20] = in_blas ? 1u : 0u;
                if (want_gen) counts = wl_generate_step(&fc, &wl, W, slot_base, counts, n_region_slots, (uint32_t)tune.wl_batch, step_ctr);
                // (re-evaluated: the generate step may have refilled the ready list)
                if (WL_N_SHADE(counts) >= 32u || (WL_N_SHADE(counts) > 0u && WL_N_READY(counts) == 0u && starving))
                    counts = wl_shade_step(&fc, instances, shade, &wl, frame_sum, W, slot_base, counts, step_ctr);
                tr.o = f3(u2f(park[0]), u2f(park[1]), u2f(park[2]));
                tr.d = f3(u2f(park[3]), u2f(park[4]), u2f(park[5]));
                tr.idir = f3(u2f(park[6]), u2f(park[7]), u2f(park[8]));
                tr.oct_inv = park[9]; tr.pow4_lo = park[10]; tr.pow4_hi = park[11];
                tmax = u2f(park[12]); ngroup = make_uint2(park[13], park[14]); tgroup = make_uint2(park[15], park[16]);
                slot = park[17]; stack.sp = (int)park[18]; cur_inst = park[19]; in_blas = park[20] != 0u;
            }
            // ---- hand ready slots to idle lanes ----
            if (WL_N_READY(counts) > 0u && idle_enough) {
                const uint32_t rank = (uint32_t)__popc(~b_ray & lt_mask);
                const uint32_t cnt = min(WL_N_READY(counts), n_idle);
                if (!WL_ACTIVE() && rank < cnt) {
                    slot = (uint32_t)W->ready[((counts >> 25) + rank) & (WL_RING - 1u)];  // oldest first
                    const uint32_t gs = slot_base + slot;
                    const float4 o = wl_ld(wl.ray_o + gs), d = wl_ld(wl.ray_d + gs), id = wl_ld(wl.ray_i + gs);
                    s_frame0[tid] = wl_ld(wl.frame0 + gs);
                    {
                        const float4 f1 = wl_ld(wl.frame1 + gs);
                        s_frame1[tid] = make_float2(f1.x, f1.y);
                    }
                    tr.o = f3(o.x, o.y, o.z);
                    tr.d = f3(d.x, d.y, d.z);
                    tr.idir = f3(id.x, id.y, id.z);
                    set_trav_octant(tr, tr.d);
                    s_hit[tid] = make_uint4(SOLB_MISS, SOLB_MISS, 0u, 0u);
                    tmax = fc.tmax;
                    ngroup = SOLB_ROOT_GROUP;
                    tgroup = make_uint2(0u, 0u);
                    stack.sp = 0;
                    in_blas = false;
                }
                counts = counts - cnt + (cnt << 25);  // n_ready -= cnt (low byte), head += cnt (top 7 bits, wraps mod WL_RING)
                b_ray = __ballot_sync(0xffffffffu, WL_ACTIVE());
                __syncwarp();
            }
            if (b_ray == 0u) {
                // every finished lane has been collected above, so the lists account for all live slots
                if ((counts & 0xffffu) == 0u && WL_N_FREE(counts) == 0u) break;  // (free is 0 once the frame's cursor is exhausted)
                continue;
            }
        }
        // ---- vote: node step or triangle step (k_wf_trace) ----
        const bool w_node = (ngroup.y & 0xff000000u) != 0u;
        const bool w_tri = tgroup.y != 0u;
        const uint32_t b_node = __ballot_sync(0xffffffffu, w_node), b_tri = __ballot_sync(0xffffffffu, w_tri);
        const int nn = __popc(b_node), nt = __popc(b_tri);
        if (nt > 0 && nt >= nn) {  // (weights other than 1 : 1 measured slower: profiles/r02_sweep_warpfront_knobs.txt)
            if (w_tri) {
                if (TL && !in_blas) {
                    trav_enter_instance(inst_leaves, tr.o, tr.d, tr, ngroup, tgroup, cur_inst, stack);
                    s_frame0[tid] = make_float4(tr.frame.e1.x, tr.frame.e1.y, tr.frame.e1.z, tr.frame.e2.x);
                    s_frame1[tid] = make_float2(tr.frame.e2.y, tr.frame.e2.z);
                    in_blas = true;
                } else {
                    const float4 f0 = s_frame0[tid];
                    const float2 f1 = s_frame1[tid];
                    tr.frame.e1 = f3(f0.x, f0.y, f0.z);
                    tr.frame.e2 = f3(f0.w, f1.x, f1.y);
                    Hit h;
                    h.gtri = s_hit[tid].y;  // the hit in hand (tie-break of equal t)
                    if (trav_tri_step(tris, tr, tmax, tgroup, h))
                        s_hit[tid] = make_uint4(TL ? cur_inst : h.inst, h.gtri, __float_as_uint(h.u), __float_as_uint(h.v));
                    if (STATS) ctr.tris++;
                }
            }
        } else if (w_node) {
            if (tgroup.y) stack.push(tgroup);  // postpone the pending triangles
            trav_node_step(nodes, tr, tmax, ngroup, tgroup, stack);
            if (STATS) ctr.nodes++;
        }
        // ---- lanes with nothing in hand: pop, or finish the ray (its slot is collected at the next service) ----
        if (!(ngroup.y & 0xff000000u) && !tgroup.y && !stack.empty()) {
            {
                const uint2 e = stack.pop();
                if (TL && e.y == 0u) {  // sentinel: back to the TLAS with the world-space ray
                    const float4 o = wl_ld(wl.ray_o + (slot_base + slot)), d = wl_ld(wl.ray_d + (slot_base + slot));
                    tr = make_trav_ray(f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), fc.tmin);
                    s_frame0[tid] = make_float4(tr.frame.e1.x, tr.frame.e1.y, tr.frame.e1.z, tr.frame.e2.x);
                    s_frame1[tid] = make_float2(tr.frame.e2.y, tr.frame.e2.z);
                    in_blas = false;
                } else if (e.y & 0xff000000u) ngroup = e;
                else tgroup = e;
            }
        }
        b_ray = __ballot_sync(0xffffffffu, WL_ACTIVE());
    }
#undef WL_ACTIVE
    warp_add_stat(stats, ST_RAYS, step_ctr[2]);
    warp_add_stat(stats, ST_HITS, step_ctr[0]);
    warp_add_stat(stats, ST_PATHS, step_ctr[1]);
    if (STATS) {
        warp_add_stat(stats, ST_NODES, ctr.nodes);
        warp_add_stat(stats, ST_TRIS, ctr.tris);
    }
}

// ---------------------------------------------------------------------------------------------------
// 4-ray-ao: ao.rgen:37-83 + ao.rchit:45-88 + ao.rmiss:7-10, one thread per pixel
// The sample loop (ao.rgen:47-75) and the chained-ray loop (ao.rgen:56-72) are flattened into one loop so that, with VOTE, the
// whole warp can step its traversals together (trace_vote); per-pixel arithmetic and RNG order are those of the nested loops.
template <bool TL, bool VOTE>
__global__ void __launch_bounds__(TRACE_BLOCK) k_ao(const FrameConsts fc, const uint4 *__restrict__ nodes,
                                                    const float4 *__restrict__ tris, const float4 *__restrict__ inst_leaves,
                                                    const DeviceInstance *__restrict__ instances,
                                                    const ShadeRecord *__restrict__ shade, const uint32_t *__restrict__ blue,
                                                    uint32_t blue_w, uint32_t blue_h, float4 *image, unsigned long long *stats) {
    SOLB_DECL_STACK();
    uint32_t x, y;
    const bool active = thread_pixel(fc, x, y);
    uint32_t nr = 0, nh = 0, np = 0;
    const uint32_t max_samples = fc.max_bounces;  // ao.rgen:42 (4)
    const uint32_t sample_count = fc.spp;         // ao.rgen:43 (4)
    uint32_t rng = 0, s = 0, depth = 0;
    float3 ao = f3(0, 0, 0), hit_value = f3(0, 0, 0);
    Ray r;
    r.o = fc.origin; r.d = f3(0, 0, 1);
    r.tmin = fc.tmin;  // preparePayload, ao.rgen:33
    r.tmax = fc.tmax;
    bool live = active && sample_count > 0u;
    if (live) {
        rng = tea(x + y * fc.width, fc.frame);                 // ao.rgen:45
        const float jx = next_rand(rng), jy = next_rand(rng);  // ao.rgen:49
        r.d = primary_dir(fc, (float)x + jx, (float)y + jy);
        np++;
    }
    while (VOTE ? __any_sync(0xffffffffu, live) : live) {
        Hit h;
        if (VOTE) {
            trace_vote<false, TL>(nodes, tris, inst_leaves, r, live, h, stack, (TraceCounters *)nullptr);
        } else {
            stack.sp = 0;
            trace_any<false, TL>(nodes, tris, inst_leaves, r, h, stack, (TraceCounters *)nullptr);
        }
        if (live) {
            nr++;
            bool end_sample = true;  // ao.rmiss: done = 1
            if (h.inst != SOLB_MISS) {
                nh++;
                ao_hit(instances, shade, h.inst, h.gtri, h.u, h.v, x, y, blue, blue_w, blue_h, depth, s, r.o, r.d, rng);
                r.tmin = 0.001f;  // ao.rchit:83
                r.tmax = 10.0f;
                if (depth > 0) hit_value = hit_value + f3(1, 1, 1);  // ao.rchit:84-86
                depth++;
                end_sample = depth > max_samples;  // ao.rgen:71
            }
            if (end_sample) {
                ao = ao + hit_value * (1.0f / (float)max_samples);  // ao.rgen:74
                s++;
                if (s < sample_count) {
                    const float jx = next_rand(rng), jy = next_rand(rng);  // ao.rgen:49
                    r.o = fc.origin;
                    r.d = primary_dir(fc, (float)x + jx, (float)y + jy);
                    r.tmin = fc.tmin;
                    r.tmax = fc.tmax;
                    hit_value = f3(0, 0, 0);
                    depth = 0;
                    np++;
                } else {
                    live = false;
                }
            }
        }
    }
    if (active) {
        float3 color = f3(1.0f - ao.x / (float)sample_count, 1.0f - ao.y / (float)sample_count, 1.0f - ao.z / (float)sample_count);
        const size_t p = (size_t)y * fc.width + x;
        const float a = 1.0f / (float)(uint32_t)(fc.frame - (uint32_t)fc.accum_start + 1u);  // ao.rgen:78
        const float4 old4 = image[p];
        color = mix3(f3(old4.x, old4.y, old4.z), color, a);  // ao.rgen:80 (no NaN guard, no gamma)
        image[p] = make_float4(color.x, color.y, color.z, 1.0f);
    }
    warp_add_stat(stats, ST_RAYS, nr);
    warp_add_stat(stats, ST_HITS, nh);
    warp_add_stat(stats, ST_PATHS, np);
}

__global__ void __launch_bounds__(256) k_resolve_sum(const float4 *sum, float4 *accum_out, uint32_t *render, uint32_t n) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float4 s = sum[p];
    const float inv = s.w > 0.0f ? 1.0f / s.w : 0.0f;
    const float3 c = f3(s.x * inv, s.y * inv, s.z * inv);
    if (accum_out) accum_out[p] = make_float4(c.x, c.y, c.z, 1.0f);
    const float g = 1.0f / 2.2f;
    if (render) render[p] = pack_rgba8(powf(c.x, g), powf(c.y, g), powf(c.z, g), 1.0f);
}

// de-index the reference vertex layout into per-triangle shading records (object space)
__global__ void k_build_shade_records(const DeviceSceneView sv, ShadeRecord *out) {
    const float4 *__restrict__ vertices = sv.vertices;  // 4 float4 per ModelVertex
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;  // geometry triangle: records are shared by all instances of a BLAS
    if (g >= sv.n_geom_tris) return;
    uint32_t lo = 0, hi = sv.n_blas;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sv.blas[mid].first_tri <= g) lo = mid; else hi = mid;
    }
    const DeviceBlas di = sv.blas[lo];
    const uint32_t prim = g - di.first_tri;
    float f[28];
    for (int k = 0; k < 3; k++) {
        const uint32_t vi = di.first_vertex + sv.indices[di.first_index + 3 * prim + k];
        const float4 pos = vertices[4 * (size_t)vi + 0], col = vertices[4 * (size_t)vi + 1], nrm = vertices[4 * (size_t)vi + 2];
        f[9 * k + 0] = pos.x; f[9 * k + 1] = pos.y; f[9 * k + 2] = pos.z;
        f[9 * k + 3] = nrm.x; f[9 * k + 4] = nrm.y; f[9 * k + 5] = nrm.z;
        f[9 * k + 6] = col.x; f[9 * k + 7] = col.y; f[9 * k + 8] = col.z;
    }
    f[27] = 0.0f;
    for (int i = 0; i < 7; i++) out[g].q[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}

// ---------------------------------------------------------------------------------------------------
// host-side launch wrappers
cudaError_t launch_build_shade_records(cudaStream_t st, const DeviceSceneView &sv, ShadeRecord *out) {
    if (sv.n_geom_tris == 0) return cudaSuccess;
    k_build_shade_records<<<(sv.n_geom_tris + 255) / 256, 256, 0, st>>>(sv, out);
    return cudaGetLastError();
}

cudaError_t launch_debug(cudaStream_t st, const FrameConsts &fc, const AccelStorage &as, uint32_t *render, uint2 *ids,
                         float4 *attribs, unsigned long long *stats, const TraceTuning &tune) {
    const uint32_t blocks = pixel_grid_blocks(fc);
    const bool vote = tune.mega_vote < 0 ? as.n_wide > 8u : tune.mega_vote != 0;
#define SOLB_DBG(T, V) k_debug<T, V><<<blocks, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), T ? as.inst_leaves_f4() : nullptr, render, ids, attribs, stats)
    if (as.two_level) { if (vote) SOLB_DBG(true, true); else SOLB_DBG(true, false); }
    else { if (vote) SOLB_DBG(false, true); else SOLB_DBG(false, false); }
#undef SOLB_DBG
    return cudaGetLastError();
}

cudaError_t launch_trace_rays(cudaStream_t st, const AccelStorage &as, const float4 *rays, uint32_t n, uint4 *hits, float *t_out,
                              unsigned long long *stats, const TraceTuning &tune) {
    if (n == 0) return cudaSuccess;
    const uint32_t blocks = (n + TRACE_BLOCK - 1) / TRACE_BLOCK;
    const bool vote = tune.mega_vote < 0 ? as.n_wide > 8u : tune.mega_vote != 0;
#define SOLB_TR(T, V) k_trace_rays<T, V><<<blocks, TRACE_BLOCK, 0, st>>>(as.nodes_u4(), as.tris_f4(), T ? as.inst_leaves_f4() : nullptr, rays, n, hits, t_out, stats)
    if (as.two_level) { if (vote) SOLB_TR(true, true); else SOLB_TR(true, false); }
    else { if (vote) SOLB_TR(false, true); else SOLB_TR(false, false); }
#undef SOLB_TR
    return cudaGetLastError();
}

cudaError_t launch_pathtrace_mega(cudaStream_t st, const FrameConsts &fc, const AccelStorage &as, const DeviceInstance *instances,
                                  const ShadeRecord *shade, float4 *accum, uint32_t *render, unsigned long long *stats,
                                  bool collect, uint32_t *pixel_counter, int sm_count, const TraceTuning &tune) {
    const uint32_t blocks = pixel_grid_blocks(fc);
    const float4 *il = as.inst_leaves_f4();
    const bool vote = tune.mega_vote < 0 ? as.n_wide > 8u : tune.mega_vote != 0;
    if (pixel_counter && tune.mega_persistent) {
        const uint32_t n_slots = ((fc.width + 7u) >> 3) * region_tiles_y(fc) * 32u;
        const uint32_t grid = std::min<uint32_t>(blocks, (uint32_t)(sm_count * tune.mega_ctas_per_sm));
        cudaError_t err = cudaMemsetAsync(pixel_counter, 0, sizeof(uint32_t), st);
        if (err != cudaSuccess) return err;
#define SOLB_MEGA_PV(S, T, V) k_pathtrace_mega_persistent<S, T, V><<<grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), T ? il : nullptr, instances, shade, accum, render, stats, pixel_counter, n_slots, tune.mega_fetch_idle)
#define SOLB_MEGA_P(S, T) do { if (vote) SOLB_MEGA_PV(S, T, true); else SOLB_MEGA_PV(S, T, false); } while (0)
        if (as.two_level) { if (collect) SOLB_MEGA_P(true, true); else SOLB_MEGA_P(false, true); }
        else { if (collect) SOLB_MEGA_P(true, false); else SOLB_MEGA_P(false, false); }
#undef SOLB_MEGA_P
#undef SOLB_MEGA_PV
        return cudaGetLastError();
    }
#define SOLB_MEGA(S, T, V) k_pathtrace_mega<S, T, V><<<blocks, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), T ? il : nullptr, instances, shade, accum, render, stats)
#define SOLB_MEGA_V(S, T) do { if (vote) SOLB_MEGA(S, T, true); else SOLB_MEGA(S, T, false); } while (0)
    if (as.two_level) { if (collect) SOLB_MEGA_V(true, true); else SOLB_MEGA_V(false, true); }
    else { if (collect) SOLB_MEGA_V(true, false); else SOLB_MEGA_V(false, false); }
#undef SOLB_MEGA_V
#undef SOLB_MEGA
    return cudaGetLastError();
}

cudaError_t launch_ao(cudaStream_t st, const FrameConsts &fc, const AccelStorage &as, const DeviceInstance *instances,
                      const ShadeRecord *shade, const uint32_t *blue, uint32_t bw, uint32_t bh, float4 *image,
                      unsigned long long *stats, const TraceTuning &tune) {
    const uint32_t blocks = pixel_grid_blocks(fc);
    const bool vote = tune.mega_vote < 0 ? as.n_wide > 8u : tune.mega_vote != 0;
#define SOLB_AO(T, V) k_ao<T, V><<<blocks, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), T ? as.inst_leaves_f4() : nullptr, instances, shade, blue, bw, bh, image, stats)
    if (as.two_level) { if (vote) SOLB_AO(true, true); else SOLB_AO(true, false); }
    else { if (vote) SOLB_AO(false, true); else SOLB_AO(false, false); }
#undef SOLB_AO
    return cudaGetLastError();
}

cudaError_t launch_resolve_sum(cudaStream_t st, const float4 *sum, float4 *accum_out, uint32_t *render, uint32_t n) {
    if (n == 0) return cudaSuccess;
    k_resolve_sum<<<(n + 255) / 256, 256, 0, st>>>(sum, accum_out, render, n);
    return cudaGetLastError();
}

// One reference frame with the wavefront schedule.
//
// The frame is split into 1..4 horizontal parts of whole tile rows.  Each part runs its own wave sequence (own
// queues + counters, shared per-pixel state arrays) on its own stream, so the drain tail of one part's persistent
// trace kernel and its memory-bound shade kernel overlap with another part's traversal.
// (body: every early return leaves through launch_pathtrace_wavefront's tail, which lets pending count copies land and joins
// the part streams back into the ctx stream, so later work on the ctx stream never races kernels still running on a part stream)
static cudaError_t wavefront_body(const WavefrontLaunch &L, const FrameConsts &fc, const AccelStorage &as,
                                  const DeviceInstance *instances, const ShadeRecord *shade, float4 *accum, uint32_t *render,
                                  unsigned long long *stats, bool collect, uint64_t *launches, std::vector<cudaEvent_t> *events,
                                  uint32_t *n_events_used, const TraceTuning &tune, int *forked_lanes, bool *pending) {
    cudaError_t err = cudaSuccess;
    const uint32_t tiles_x = (fc.width + 7u) >> 3, tiles_y = region_tiles_y(fc);
    int n_lanes = (events || !L.stream[1]) ? 1 : tune.overlap;
    if (n_lanes > WF_MAX_PARTS) n_lanes = WF_MAX_PARTS;
    {   // small launches (a rank's share of a tile-split frame, small images): a part needs enough rays to fill its persistent grid
        const uint64_t region_pixels = (uint64_t)fc.width * fc.band_rows * fc.n_bands;
        const int by_size = (int)(region_pixels / (uint64_t)tune.min_pixels_per_part);
        if (n_lanes > 1 && by_size < n_lanes) n_lanes = by_size < 1 ? 1 : by_size;
    }
    if ((uint32_t)n_lanes > tiles_y) n_lanes = (int)tiles_y;
    if (n_lanes < 1) n_lanes = 1;
    const uint32_t rows_per_lane = (tiles_y + n_lanes - 1) / n_lanes;
    const int trace_grid = L.sm_count * (n_lanes > 1 ? tune.ctas_per_sm_overlap : tune.ctas_per_sm);  // persistent CTAs of 128 threads
    const int shade_grid = L.sm_count * (n_lanes > 1 ? tune.shade_ctas_per_sm_overlap : 4);          // grid-stride
    // every pixel needs at least spp rays; at most spp * (max_bounces + 2)
    const uint32_t max_waves = fc.spp * (fc.max_bounces + 2u);
    const uint32_t check_every = (uint32_t)tune.check_every;
    uint32_t part_slot_begin[WF_MAX_PARTS], part_slots[WF_MAX_PARTS];
    bool live[WF_MAX_PARTS] = { false, false, false, false };
    int qi[WF_MAX_PARTS] = { 0, 0, 0, 0 };
    if (n_lanes > 1) {  // fork: the extra streams start after everything already queued on the ctx stream
        if ((err = cudaEventRecord(L.fork, L.stream[0])) != cudaSuccess) return err;
        for (int k = 1; k < n_lanes; k++) {
            if ((err = cudaStreamWaitEvent(L.stream[k], L.fork, 0)) != cudaSuccess) return err;
            *forked_lanes = k + 1;
        }
    }
    for (int k = 0; k < n_lanes; k++) {
        const uint32_t row0 = k * rows_per_lane, row1 = (row0 + rows_per_lane < tiles_y) ? row0 + rows_per_lane : tiles_y;
        const uint32_t slot_begin = row0 * tiles_x * 32u, n_slots = (row1 - row0) * tiles_x * 32u;
        part_slot_begin[k] = slot_begin;
        part_slots[k] = n_slots;
        if ((err = cudaMemsetAsync(L.ws[k].counters, 0, 4 * sizeof(uint32_t), L.stream[k])) != cudaSuccess) return err;
        k_wf_generate<<<(n_slots + 255) / 256, 256, 0, L.stream[k]>>>(fc, L.ws[k], slot_begin, n_slots, stats);
        *launches += 1;
        live[k] = true;
    }
    // Every part advances through its own wave sequence.  The host stays up to tune.max_ahead waves ahead of the last survivor
    // count it has SEEN for a part (async_poll: a count is copied back every check_every waves and picked up with a non-blocking
    // event query), so parts never wait for each other on the host and a part on a higher-priority stream can run ahead of the
    // others: its drain tail then overlaps their full waves instead of their tails.  async_poll = 0 restores the lock-step
    // schedule (all parts synchronised every check_every waves).
    uint32_t wave_k[WF_MAX_PARTS] = { 0, 0, 0, 0 }, confirmed[WF_MAX_PARTS] = { 0, 0, 0, 0 }, poll_wave[WF_MAX_PARTS] = { 0, 0, 0, 0 };
    const bool async_poll = tune.async_poll && !events && L.poll[0];
    const uint32_t max_ahead = (uint32_t)std::max(tune.max_ahead, tune.check_every);
    auto launch_wave = [&](int k) {
        cudaStream_t st = L.stream[k];
        if (events) {  // timing mode (single lane): bracket the dominant kernel with an event pair
            while (events->size() < (size_t)*n_events_used + 2) {
                cudaEvent_t e;
                if ((err = cudaEventCreate(&e)) != cudaSuccess) return;
                events->push_back(e);
            }
            cudaEventRecord((*events)[*n_events_used], st);
        }
        if (as.two_level) {
            const float4 *il = as.inst_leaves_f4();
            if (collect) k_wf_trace<true, true, false><<<trace_grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), il, L.ws[k], qi[k], stats, tune);
            else k_wf_trace<false, true, false><<<trace_grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), il, L.ws[k], qi[k], stats, tune);
        } else if (tune.pool && L.spill[k]) {
            const int pool_grid = L.sm_count * tune.pool_ctas_per_sm;
            if (collect) k_wf_trace_pool<true><<<pool_grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), L.ws[k], qi[k], stats, tune, L.spill[k]);
            else k_wf_trace_pool<false><<<pool_grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), L.ws[k], qi[k], stats, tune, L.spill[k]);
        } else if (tune.coop_tri && as.n_tris < (1u << 27)) {  // the dealt item packs the triangle index into 27 bits
            if (collect) k_wf_trace<true, false, true><<<trace_grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), nullptr, L.ws[k], qi[k], stats, tune);
            else k_wf_trace<false, false, true><<<trace_grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), nullptr, L.ws[k], qi[k], stats, tune);
        } else if (collect) k_wf_trace<true, false, false><<<trace_grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), nullptr, L.ws[k], qi[k], stats, tune);
        else k_wf_trace<false, false, false><<<trace_grid, TRACE_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), nullptr, L.ws[k], qi[k], stats, tune);
        if (events) {
            cudaEventRecord((*events)[*n_events_used + 1], st);
            *n_events_used += 2;
        }
        // shade_priority: the short, memory-bound shade kernel runs on a high-priority side stream so its blocks are placed as
        // soon as any persistent trace CTA of another part retires, instead of queueing behind those parts' pending CTAs
        cudaStream_t ss = (tune.shade_priority && !events && L.shade_stream[k]) ? L.shade_stream[k] : st;
        if (ss != st) {
            cudaEventRecord(L.ev_ts[k], st);
            cudaStreamWaitEvent(ss, L.ev_ts[k], 0);
        }
        if (tune.sort_shade) k_wf_shade<true><<<shade_grid, 256, 0, ss>>>(fc, instances, shade, L.ws[k], qi[k], stats);
        else k_wf_shade<false><<<shade_grid, 256, 0, ss>>>(fc, instances, shade, L.ws[k], qi[k], stats);
        if (ss != st) {
            cudaEventRecord(L.ev_st[k], ss);
            cudaStreamWaitEvent(st, L.ev_st[k], 0);
        }
        *launches += 2;
        qi[k] ^= 1;
        wave_k[k]++;
    };
    if (!async_poll) {
        for (uint32_t wave = 0; wave < max_waves && (live[0] || live[1] || live[2] || live[3]); wave++) {
            for (int k = 0; k < n_lanes; k++) {
                if (!live[k]) continue;
                launch_wave(k);
                if (err != cudaSuccess) return err;
            }
            if (wave + 1 >= fc.spp && ((wave + 1) % check_every) == 0) {
                // poll the survivor counts so finished parts stop launching empty waves
                for (int k = 0; k < n_lanes; k++)
                    if (live[k] && (err = cudaMemcpyAsync(&L.host_counts[k], &L.ws[k].counters[qi[k]], sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                                          L.stream[k])) != cudaSuccess) return err;
                for (int k = 0; k < n_lanes; k++) {
                    if (!live[k]) continue;
                    if ((err = cudaStreamSynchronize(L.stream[k])) != cudaSuccess) return err;
                    if (L.host_counts[k] == 0) live[k] = false;
                }
            }
        }
    } else {
        for (;;) {
            bool any = false;
            for (int k = 0; k < n_lanes; k++) {
                if (!live[k]) continue;
                if (wave_k[k] >= max_waves) { live[k] = false; continue; }
                if (pending[k]) {
                    // too far ahead of the last count seen: wait for the outstanding one
                    cudaError_t q = (wave_k[k] - confirmed[k] >= max_ahead) ? cudaEventSynchronize(L.poll[k]) : cudaEventQuery(L.poll[k]);
                    if (q == cudaSuccess) {
                        pending[k] = false;
                        confirmed[k] = poll_wave[k];
                        if (L.host_counts[k] == 0) { live[k] = false; continue; }
                    } else if (q != cudaErrorNotReady) return q;
                }
                any = true;
                launch_wave(k);
                if (err != cudaSuccess) return err;
                if (!pending[k] && wave_k[k] >= fc.spp && wave_k[k] - confirmed[k] >= check_every) {
                    if ((err = cudaMemcpyAsync(&L.host_counts[k], &L.ws[k].counters[qi[k]], sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                               L.stream[k])) != cudaSuccess) return err;
                    if ((err = cudaEventRecord(L.poll[k], L.stream[k])) != cudaSuccess) return err;
                    pending[k] = true;
                    poll_wave[k] = wave_k[k];
                }
            }
            if (!any) break;
        }
    }
    for (int k = 0; k < n_lanes; k++) {
        k_wf_resolve<<<(part_slots[k] + 255) / 256, 256, 0, L.stream[k]>>>(fc, L.ws[k], accum, render, part_slot_begin[k], part_slots[k]);
        *launches += 1;
    }
    return cudaGetLastError();
}

cudaError_t launch_pathtrace_wavefront(const WavefrontLaunch &L, const FrameConsts &fc, const AccelStorage &as,
                                       const DeviceInstance *instances, const ShadeRecord *shade, float4 *accum, uint32_t *render,
                                       unsigned long long *stats, bool collect, uint64_t *launches, std::vector<cudaEvent_t> *events,
                                       uint32_t *n_events_used, const TraceTuning &tune) {
    if (fc.width == 0 || fc.band_rows == 0 || fc.n_bands == 0) return cudaSuccess;
    int forked_lanes = 1;
    bool pending[WF_MAX_PARTS] = { false, false, false, false };
    cudaError_t err = wavefront_body(L, fc, as, instances, shade, accum, render, stats, collect, launches, events, n_events_used, tune,
                                     &forked_lanes, pending);
    // common tail, also after an error: a count still in flight targets pinned memory (let it land before the next call reuses
    // the slot), and every forked part stream is joined back into the ctx stream
    for (int k = 0; k < WF_MAX_PARTS; k++)
        if (pending[k]) {
            const cudaError_t e = cudaEventSynchronize(L.poll[k]);
            if (err == cudaSuccess) err = e;
        }
    for (int k = 1; k < forked_lanes; k++) {
        cudaError_t e = cudaEventRecord(L.join[k], L.stream[k]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(L.stream[0], L.join[k], 0);
        if (err == cudaSuccess) err = e;
    }
    return err;
}

uint32_t warpfront_grid_warps(int sm_count, const TraceTuning &tune) {
    return (uint32_t)(sm_count * (tune.wl_warps_per_sm > 0 ? tune.wl_warps_per_sm : tune.wl_ctas_per_sm * (TRACE_BLOCK / 32)));
}

// One reference frame with the warp-local wavefront schedule: a cursor reset and ONE launch on `st` (a side stream of the ctx:
// the kernel writes only its own path state and frame_sum, the per-pixel sums of the frame's samples) ...
cudaError_t launch_pathtrace_warpfront(cudaStream_t st, const FrameConsts &fc, const AccelStorage &as, const DeviceInstance *instances,
                                       const ShadeRecord *shade, const WarpfrontState &wl, float4 *frame_sum,
                                       unsigned long long *stats, bool collect, int sm_count, const TraceTuning &tune) {
    if (fc.width == 0 || fc.band_rows == 0 || fc.n_bands == 0) return cudaSuccess;
    const uint32_t n_slots = ((fc.width + 7u) >> 3) * region_tiles_y(fc) * 32u;
    cudaError_t err = cudaMemsetAsync(wl.cursor, 0, sizeof(uint32_t), st);
    if (err != cudaSuccess) return err;
    // small regions (a rank's share of a tile-split frame, tiny images): fewer warps rather than smaller pools.  A warp needs
    // ~72 slots to keep 32 lanes tracing while 32 finished rays wait for a full-width shade step; with every resident warp
    // holding 56 pixels (1/8 of a 1080p frame over 8 CTAs per SM) the shade steps ran partial and the share took 6.2 ms, with
    // 6 CTAs per SM (72 slots each) 4.9 ms (tools/tile_time.py, profiles/r02_tile_split.txt).
    const uint32_t want_warps = std::min<uint32_t>(std::min(warpfront_grid_warps(sm_count, tune), wl.n_warps), std::max<uint32_t>(1u, n_slots / (uint32_t)std::max(tune.wl_region_slots_per_warp, 32)));
    const uint32_t grid = (want_warps + WL_BLOCK / 32 - 1) / (WL_BLOCK / 32);
    // pixels a warp holds at a time: its even share of the region, between one warp-width and the full pool
    const uint32_t share = (n_slots + want_warps - 1) / want_warps;
    const uint32_t pool_limit = std::min<uint32_t>((uint32_t)WL_POOL, std::max<uint32_t>(32u, (share + 7u) & ~7u));
    const float4 *il = as.inst_leaves_f4();
    TraceTuning tn = tune;
    if ((size_t)as.n_wide * sizeof(Node8) + (size_t)as.n_tris * sizeof(Tri48) > ((size_t)64 << 20)) tn.wl_fetch_idle = tune.wl_fetch_idle_large;
#define SOLB_WLF(S, T) k_pt_warpfront<S, T><<<grid, WL_BLOCK, 0, st>>>(fc, as.nodes_u4(), as.tris_f4(), T ? il : nullptr, instances, shade, wl, frame_sum, stats, n_slots, pool_limit, tn)
    if (as.two_level) { if (collect) SOLB_WLF(true, true); else SOLB_WLF(false, true); }
    else { if (collect) SOLB_WLF(true, false); else SOLB_WLF(false, false); }
#undef SOLB_WLF
    return cudaGetLastError();
}

// ... and its resolve on the ctx stream: pathtrace.rgen:88-103 per pixel of the launch region (k_wf_resolve reads pix[p].xyz).
cudaError_t launch_warpfront_resolve(cudaStream_t st, const FrameConsts &fc, float4 *frame_sum, float4 *accum, uint32_t *render) {
    if (fc.width == 0 || fc.band_rows == 0 || fc.n_bands == 0) return cudaSuccess;
    const uint32_t n_slots = ((fc.width + 7u) >> 3) * region_tiles_y(fc) * 32u;
    WavefrontState ws = {};
    ws.pix = frame_sum;
    k_wf_resolve<<<(n_slots + 255) / 256, 256, 0, st>>>(fc, ws, accum, render, 0u, n_slots);
    return cudaGetLastError();
}

size_t pool_spill_bytes(int sm_count, const TraceTuning &tune) {
    return (size_t)sm_count * tune.pool_ctas_per_sm * PL_WARPS * PL_SLOTS * PL_SPILL * sizeof(uint2);
}

}  // namespace solb
