// trace.h — host-callable launch wrappers of trace.cu
#pragma once
#include <vector>

#include "solb_internal.h"

namespace solb {

// Wavefront path state, one slot per pixel (SoA, 16-byte vector loads/stores):
struct WavefrontState {
    float4 *ray_o;      // xyz origin of the ray in flight
    float4 *ray_d;      // xyz direction
    float4 *thr;        // xyz throughput (accumulatedRayColor), w = bits(depth | sample << 16)
    float4 *pix;        // xyz sum of finished samples (pixelColor), w = bits(prd.rng)
    uint4 *hit;         // instance, global triangle, bits(u), bits(v)
    uint32_t *queue[2]; // pixel ids with a ray in flight (ping-pong)
    uint32_t *counters; // [0],[1]: queue sizes, [2]: trace work head
    uint32_t capacity;  // pixels
};

// scheduling knobs of the wavefront trace kernel (defaults chosen by measurement, profiles/; overridable through
// SOLB_FETCH_IDLE / SOLB_TRI_WEIGHT / SOLB_NODE_WEIGHT / SOLB_CTAS_PER_SM / SOLB_CHECK_EVERY for experiments)
struct TraceTuning {
    int fetch_idle = 8;    // refill terminated lanes once this many are idle
    int tri_weight = 1;    // triangle step wins the vote when n_tri * tri_weight >= n_node * node_weight
    int node_weight = 1;
    int ctas_per_sm = 8;   // persistent grid = SMs x this
    int min_batch = 32;    // smallest ray batch a warp takes from the queue (short queues: rays / resident warps, rounded to 32)
    int check_every = 8;   // host polls the survivor count every this many waves
    int async_poll = 1;    // 1: counts are picked up with non-blocking event queries, parts advance independently on the host
    int max_ahead = 64;    // ... but never more than this many waves beyond the last count seen
    int shade_priority = 0; // 1: shade kernels on high-priority side streams (two event hops per wave): 3 453 -> 3 475 Mrays/s,
                            // inside run-to-run noise (profiles/r01_sweep_async_poll.txt): off
    int part_priority = 0; // 1: the extra part streams get increasing CUDA stream priorities (part k above part k-1) so the parts
                           // drift apart and one part's drain tail meets the others' full waves.  Measured slower (3 347 ->
                           // 3 286 Mrays/s, profiles/r01_sweep_async_poll.txt): off
    int overlap = 3;       // frame parts (1..4) run as independent wave sequences on their own streams, so the drain tail of one
                           // part's persistent trace kernel and its memory-bound shade kernel overlap another part's traversal
    int ctas_per_sm_overlap = 5;  // persistent CTAs per SM and part when overlapping
    int min_pixels_per_part = 1;  // fewer parts when a part would hold fewer pixels than this (1: always `overlap` parts)
    int coop_tri = 0;      // 1: cooperative triangle step of k_wf_trace (flattened scenes) instead of every lane testing its own
                           // triangles.  Measured slower (3 051 vs 3 300 Mrays/s): dealing the tests out costs ~170 instructions per
                           // round on top of the ~180 of the test, the policy replay only pays off below ~40
                           // (profiles/r01_sweep_coop_tri.txt).  Off by default.
    int coop_block = 8;    // ... taken once this many lanes hold only triangle work (or no lane has node work)
    int shade_ctas_per_sm_overlap = 2;  // grid-stride shade CTAs (256 threads) per SM and part when overlapping
    int sort_shade = 0;        // 1: material-sorted shading (block-level counting sort of the shade queue by hit class)
    int pool = 0;              // 1: ray-pool traversal kernel (k_wf_trace_pool) instead of the lane-bound k_wf_trace
    int pool_ctas_per_sm = 6;  // its persistent CTAs per SM (34 KB shared memory each)
    int pool_refill = 16;      // refill free slots once this many of a warp's 64 are free
    int mega_persistent = 0;   // megakernel schedule: persistent CTAs with dynamic pixel fetch instead of one thread per pixel.
                               // Measured slower (cornell 1080p 8 027 vs 8 459 Mrays/s, tunnel 2 236 vs 2 284): the hardware block
                               // scheduler already refills finished warps, only the intra-warp imbalance is left to win and the
                               // refill code costs more than that.  Off by default (SOLB_MEGA_PERSISTENT=1 to compare).
    int mega_vote = -1;        // megakernel schedule: warp-voted traversal (trace_vote) instead of every lane running its own
                               // node-step / triangle-loop sequence (trace_closest).  -1 = by hierarchy size: tunnel.gltf
                               // 2 426 -> 3 050 Mrays/s with the vote, cornell.gltf (3 wide nodes, nothing to diverge on)
                               // 9 001 -> 8 025 (gpurun_out/mega_vote.log), so the vote is used above 8 wide nodes
    int mega_ctas_per_sm = 6;  // its persistent CTAs per SM (80 registers -> 6 x 128 threads)
    int mega_fetch_idle = 8;   // refill finished lanes once this many are idle
    // warp-local wavefront schedule (k_pt_warpfront)
    int wl_ctas_per_sm = 8;    // persistent warps per SM in units of four (64 registers -> 8 x 128 threads)
    int wl_region_slots_per_warp = 72;  // small regions: launch n_slots / this many warps (see launch_pathtrace_warpfront)
    int wl_warps_per_sm = 0;   // > 0: persistent warps per SM, overrides wl_ctas_per_sm (one warp per CTA since round 2)
    int wl_fetch_idle = 20;    // hand ready rays to idle lanes once this many lanes are idle (round-2 final build, frames in flight:
                               // 16: 4 508, 20: 4 539, 24: 4 470 Mrays/s, profiles/r02_sweep_warpfront_knobs.txt)
    int wl_fetch_idle_large = 8;  // ... when the hierarchy does not fit the caches (> 64 MB of nodes + triangles): the traversal waits on
                                  // memory, not on issue slots, and more rays in flight win: 20 M-triangle scene 1 557 (20) / 1 583 (12) /
                                  // 1 605 (8) Mrays/s, profiles/r02_sweep_warpfront_knobs.txt
    int wl_starve_idle = 16;   // partial (< 32 slots) shade / generate steps only once this many lanes are idle and nothing is ready
    int wl_gen_min = 32;       // start new pixels once this many of a warp's slots are free (or its lanes starve): a full-width generate step
    int wl_frames_in_flight = 3;  // > 1: the kernel of frame f + 1 starts while frame f drains (they share nothing: each writes its
                                  // own path state and frame sums, the resolves into the targets run in order on the ctx stream); 1: serial.
                                  // Full 1080p frames: 2 and 3 give the same 4 484 Mrays/s; a rank's 1/8 share of a tile-split frame
                                  // takes 4.70 / 3.43 / 3.32 ms with 1 / 2 / 3 (profiles/r02_frames_in_flight.txt)
    int wl_batch = 32;         // pixels a warp takes from the frame's cursor per atomic (one 8x4 tile)
};

// Warp-local wavefront state: every warp of the persistent grid owns WL_POOL pixel slots, indexed warp * WL_POOL + slot.
// Same per-path record as WavefrontState, but sized by the grid (148 SMs x 8 CTAs x 4 warps x 96 slots x 80 B = 36 MB), not
// by the image, so it stays L2-resident whatever the resolution.
#ifndef SOLB_WL_POOL
#define SOLB_WL_POOL 96
#endif
constexpr int WL_POOL = SOLB_WL_POOL;
constexpr int WL_MAX_FRAMES = 3;  // frame slots a ctx keeps (path state + frame sums + side stream each)
struct WarpfrontState {
    float4 *ray_o;      // xyz origin of the ray in flight, w = bits(pixel id)
    float4 *ray_d;      // xyz direction
    // traversal set-up of that ray, computed where the ray is born (the shade / generate steps run 32 lanes wide and converged)
    // instead of by whichever lanes happen to be idle at refill time: make_trav_ray is 91 instructions and ran 17 lanes wide
    float4 *ray_i;      // xyz reciprocal direction (safe_rcp_dir)
    float4 *frame0;     // ray-space frame of the watertight test: e1.xyz, e2.x
    float4 *frame1;     // e2.y, e2.z
    float4 *thr;        // xyz throughput, w = bits(depth | sample << 16)
    float4 *pix;        // xyz sum of finished samples, w = bits(prd.rng)
    uint4 *hit;         // instance, global triangle, bits(u), bits(v)
    uint32_t *cursor;   // next pixel slot of the frame (tile order), one counter per launch
    uint32_t n_warps;   // capacity in warps
};

constexpr int WF_MAX_PARTS = 4;
struct WavefrontLaunch {
    cudaStream_t stream[WF_MAX_PARTS];   // [0] = the ctx stream; others may be null: no overlap
    cudaEvent_t fork, join[WF_MAX_PARTS];
    cudaEvent_t poll[WF_MAX_PARTS];      // survivor-count copies in flight (null: blocking polls)
    cudaStream_t shade_stream[WF_MAX_PARTS];  // high-priority side streams for the shade kernels (null: same stream as the trace)
    cudaEvent_t ev_ts[WF_MAX_PARTS], ev_st[WF_MAX_PARTS];  // trace -> shade, shade -> trace
    WavefrontState ws[WF_MAX_PARTS];     // same per-pixel state arrays, separate queues + counters
    uint2 *spill[WF_MAX_PARTS];          // per-part global stack spill of the ray-pool kernel (null: kernel unavailable)
    uint32_t *host_counts;               // pinned, WF_MAX_PARTS entries
    int sm_count;
};

cudaError_t launch_build_shade_records(cudaStream_t st, const DeviceSceneView &sv, ShadeRecord *out);
cudaError_t launch_debug(cudaStream_t st, const FrameConsts &fc, const AccelStorage &as, uint32_t *render, uint2 *ids,
                         float4 *attribs, unsigned long long *stats, const TraceTuning &tune);
cudaError_t launch_trace_rays(cudaStream_t st, const AccelStorage &as, const float4 *rays, uint32_t n, uint4 *hits, float *t_out,
                              unsigned long long *stats, const TraceTuning &tune);
cudaError_t launch_pathtrace_mega(cudaStream_t st, const FrameConsts &fc, const AccelStorage &as, const DeviceInstance *instances,
                                  const ShadeRecord *shade, float4 *accum, uint32_t *render, unsigned long long *stats,
                                  bool collect, uint32_t *pixel_counter, int sm_count, const TraceTuning &tune);
cudaError_t launch_pathtrace_wavefront(const WavefrontLaunch &L, const FrameConsts &fc, const AccelStorage &as,
                                       const DeviceInstance *instances, const ShadeRecord *shade, float4 *accum, uint32_t *render,
                                       unsigned long long *stats, bool collect, uint64_t *launches, std::vector<cudaEvent_t> *events,
                                       uint32_t *n_events_used, const TraceTuning &tune);
cudaError_t launch_ao(cudaStream_t st, const FrameConsts &fc, const AccelStorage &as, const DeviceInstance *instances,
                      const ShadeRecord *shade, const uint32_t *blue, uint32_t bw, uint32_t bh, float4 *image,
                      unsigned long long *stats, const TraceTuning &tune);
cudaError_t launch_pathtrace_warpfront(cudaStream_t st, const FrameConsts &fc, const AccelStorage &as, const DeviceInstance *instances,
                                       const ShadeRecord *shade, const WarpfrontState &wl, float4 *frame_sum,
                                       unsigned long long *stats, bool collect, int sm_count, const TraceTuning &tune);
cudaError_t launch_warpfront_resolve(cudaStream_t st, const FrameConsts &fc, float4 *frame_sum, float4 *accum, uint32_t *render);
uint32_t warpfront_grid_warps(int sm_count, const TraceTuning &tune);
size_t pool_spill_bytes(int sm_count, const TraceTuning &tune);
cudaError_t launch_resolve_sum(cudaStream_t st, const float4 *sum, float4 *accum_out, uint32_t *render, uint32_t n);

}  // namespace solb
