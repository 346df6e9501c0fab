// build.cu — GPU acceleration-structure build for sm_100a.
//
// Replaces BLAS::new / TLAS::new / TLAS::regenerate (src/ray/acceleration.rs:136-239,344-467):
//   world-space triangle setup -> 63-bit Morton codes -> LSD radix sort (8-bit digits, warp
//   match-any ranking) -> Karras hierarchy -> bottom-up refit fused with treelet SAH restructuring
//   -> level-synchronous collapse into 80-byte 8-wide nodes + 48-byte triangles.
// This TU is compiled with -Xptxas -dlcm=cg: the bottom-up passes read nodes written by other SMs in
// the same launch, so global loads must not be served from a stale L1 line.
#include "build.cuh"
#include "solb_internal.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace solb {

// ---------------------------------------------------------------------------------------------------
// ordered-uint encoding so float min/max can use integer atomics
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// bounds[0..2] = min centroid (ordered), bounds[3..5] = max centroid (ordered)
__global__ void k_init_bounds(uint32_t *bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

// One thread per global triangle g: fetch the section-relative indices, transform the three
// positions object->world with the instance matrix (SceneInstance.transform, src/ray/mod.rs:22;
// the driver would apply InstanceDescriptor.transform, acceleration.rs:330-331), emit the 48-byte
// record, the primitive box and the centroid bounds.
__global__ void k_prep_tris(const DeviceSceneView sv, Tri48 *tri_world, float4 *prim_lo, float4 *prim_hi, uint32_t *bounds) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    float3 c = f3(0, 0, 0);
    const bool valid = g < sv.n_tris;
    if (valid) {
        // binary search: last instance with first_tri <= g
        uint32_t lo = 0, hi = sv.n_instances;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sv.inst_first_tri[mid] <= g) lo = mid; else hi = mid;
        }
        const uint32_t inst = lo;
        const uint32_t prim = g - sv.inst_first_tri[inst];
        const DeviceInstance di = sv.instances[inst];
        float3 p[3];
        for (int k = 0; k < 3; k++) {
            const uint32_t vi = di.first_vertex + sv.indices[di.first_index + 3 * prim + k];
            const float4 pos = sv.vertices[4 * (size_t)vi];
            p[k] = mat4_mul_point(di.transform, f3(pos.x, pos.y, pos.z));
        }
        Tri48 t;
        t.v0 = make_float4(p[0].x, p[0].y, p[0].z, __uint_as_float(inst));
        t.v1 = make_float4(p[1].x, p[1].y, p[1].z, __uint_as_float(prim));
        t.v2 = make_float4(p[2].x, p[2].y, p[2].z, __uint_as_float(di.shade_first_tri + prim));  // shading record (geometry triangle)
        tri_world[g] = t;
        const float3 lo3 = fmin3(p[0], fmin3(p[1], p[2])), hi3 = fmax3(p[0], fmax3(p[1], p[2]));
        prim_lo[g] = make_float4(lo3.x, lo3.y, lo3.z, 0.0f);
        prim_hi[g] = make_float4(hi3.x, hi3.y, hi3.z, 0.0f);
        c = (lo3 + hi3) * 0.5f;
    }
    // warp-reduce centroid bounds, one atomic per warp per component
    float mn[3] = { valid ? c.x : 3.4e38f, valid ? c.y : 3.4e38f, valid ? c.z : 3.4e38f };
    float mx[3] = { valid ? c.x : -3.4e38f, valid ? c.y : -3.4e38f, valid ? c.z : -3.4e38f };
    for (int off = 16; off; off >>= 1)
        for (int k = 0; k < 3; k++) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], off));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
        }
    if ((threadIdx.x & 31) == 0 && mn[0] <= mx[0])
        for (int k = 0; k < 3; k++) {
            atomicMin(&bounds[k], float_to_ordered(mn[k]));
            atomicMax(&bounds[3 + k], float_to_ordered(mx[k]));
        }
}

__global__ void k_morton(const float4 *prim_lo, const float4 *prim_hi, uint32_t n, const uint32_t *bounds, uint64_t *keys,
                         uint32_t *vals) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const float3 lo = f3(ordered_to_float(bounds[0]), ordered_to_float(bounds[1]), ordered_to_float(bounds[2]));
    const float3 hi = f3(ordered_to_float(bounds[3]), ordered_to_float(bounds[4]), ordered_to_float(bounds[5]));
    const float3 ext = hi - lo;
    const float3 inv = f3(ext.x > 0.0f ? 1.0f / ext.x : 0.0f, ext.y > 0.0f ? 1.0f / ext.y : 0.0f, ext.z > 0.0f ? 1.0f / ext.z : 0.0f);
    const float4 a = prim_lo[g], b = prim_hi[g];
    const float3 c = f3((a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f, (a.z + b.z) * 0.5f);
    keys[g] = morton63(c, lo, inv);
    vals[g] = g;
}

// ---------------------------------------------------------------------------------------------------
// Onesweep LSD radix sort (Adinets & Merrill 2022), 8-bit digits, (u64 key, u32 value).
//   k_os_hist : ONE pass over the keys builds the digit histograms of all 8 passes
//   k_os_scan : exclusive scan of each pass's 256 bins -> global digit bases
//   k_os_pass : one kernel per digit pass; each tile ranks its keys (warp match-any multi-split), then resolves
//               its per-digit exclusive prefix over all earlier tiles with a chained decoupled look-back instead
//               of a separate scan kernel: keys are read once and written once per pass.
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_PASSES = 8;
constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_INC = 2u << 30, OS_VALUE = (1u << 30) - 1u;

__global__ void __launch_bounds__(RS_THREADS) k_os_hist(const uint64_t *keys, uint32_t n, uint32_t *hist /*[8][256]*/) {
    __shared__ uint32_t h[RS_PASSES][256];
    for (int i = threadIdx.x; i < RS_PASSES * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += gridDim.x * RS_THREADS) {
        const uint64_t k = keys[i];
#pragma unroll
        for (int p = 0; p < RS_PASSES; p++) atomicAdd(&h[p][(uint32_t)(k >> (8 * p)) & 0xffu], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RS_PASSES * 256; i += RS_THREADS) {
        const uint32_t v = (&h[0][0])[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// block p scans the 256 bins of pass p (exclusive), in place
__global__ void __launch_bounds__(256) k_os_scan(uint32_t *hist) {
    __shared__ uint32_t s[256];
    uint32_t *h = hist + blockIdx.x * 256;
    const uint32_t v = h[threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const uint32_t t = threadIdx.x >= (uint32_t)off ? s[threadIdx.x - off] : 0u;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    h[threadIdx.x] = s[threadIdx.x] - v;
}

__global__ void __launch_bounds__(RS_THREADS) k_os_pass(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                                                        uint32_t *vals_out, uint32_t n, int shift, const uint32_t *digit_base,
                                                        uint32_t *status /*[tiles][256], zero*/, uint32_t *tile_counter) {
    __shared__ uint32_t warp_hist[RS_WARPS][256];
    __shared__ uint32_t digit_off[256];
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);  // tile ids in start order: look-back never waits on an unstarted tile
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t chunk_base = tile * RS_TILE + warp * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t val[RS_ITEMS], rank[RS_ITEMS], dig[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const uint32_t idx = chunk_base + j * 32 + lane;
        const bool valid = idx < n;
        key[j] = valid ? keys_in[idx] : ~0ull;
        val[j] = valid ? vals_in[idx] : 0u;
        const uint32_t d = valid ? ((uint32_t)(key[j] >> shift) & 0xffu) : 0x100u;  // 0x100: padding lanes
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && (int)lane == leader) {
            old = warp_hist[warp][d];
            warp_hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = old + __popc(peers & lt_mask);
        dig[j] = d;
        __syncwarp();
    }
    __syncthreads();
    {
        const uint32_t d = threadIdx.x;
        uint32_t count = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = count;
            count += c;
        }
        // chained scan with decoupled look-back over earlier tiles, one chain per digit
        volatile uint32_t *st = status;
        uint32_t excl = 0;
        if (tile == 0) {
            st[d] = count | OS_FLAG_INC;
        } else {
            st[(size_t)tile * 256 + d] = count | OS_FLAG_AGG;
            __threadfence();
            for (int64_t t = (int64_t)tile - 1;;) {
                const uint32_t v = st[(size_t)t * 256 + d];
                const uint32_t f = v & ~OS_VALUE;
                if (f == OS_FLAG_INC) { excl += v & OS_VALUE; break; }
                if (f == OS_FLAG_AGG) { excl += v & OS_VALUE; t--; }
                // else: that tile has not published yet - spin
            }
            st[(size_t)tile * 256 + d] = ((excl + count) & OS_VALUE) | OS_FLAG_INC;
        }
        __threadfence();
        digit_off[d] = digit_base[d] + excl;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        if (dig[j] < 0x100u) {
            const uint32_t pos = digit_off[dig[j]] + warp_hist[warp][dig[j]] + rank[j];
            keys_out[pos] = key[j];
            vals_out[pos] = val[j];
        }
    }
}

// sorts (keys, vals) in place; tmp buffers hold n elements each; scratch holds onesweep_scratch_words(n) u32.
static size_t onesweep_scratch_words(uint32_t n) {
    const size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    return RS_PASSES * 256 + 8 + tiles * 256;
}
static cudaError_t radix_sort_pairs(cudaStream_t st, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp,
                                    uint32_t n, int key_bits, uint32_t *scratch, uint64_t *launches) {
    const uint32_t num_tiles = (n + RS_TILE - 1) / RS_TILE;
    uint32_t *hist = scratch, *tile_counter = scratch + RS_PASSES * 256, *status = scratch + RS_PASSES * 256 + 8;
    cudaError_t err = cudaMemsetAsync(scratch, 0, (RS_PASSES * 256 + 8) * sizeof(uint32_t), st);
    if (err != cudaSuccess) return err;
    const uint32_t hist_blocks = std::min<uint32_t>(num_tiles, 148u * 8u);
    k_os_hist<<<hist_blocks, RS_THREADS, 0, st>>>(keys, n, hist);
    k_os_scan<<<RS_PASSES, 256, 0, st>>>(hist);
    *launches += 2;
    uint64_t *kin = keys, *kout = keys_tmp;
    uint32_t *vin = vals, *vout = vals_tmp;
    int passes = (key_bits + 7) / 8;
    if (passes & 1) passes++;  // even number of passes so the result lands back in (keys, vals)
    if (passes > RS_PASSES) passes = RS_PASSES;
    for (int p = 0; p < passes; p++) {
        if ((err = cudaMemsetAsync(status, 0, (size_t)num_tiles * 256 * sizeof(uint32_t), st)) != cudaSuccess) return err;
        if ((err = cudaMemsetAsync(tile_counter, 0, sizeof(uint32_t), st)) != cudaSuccess) return err;
        k_os_pass<<<num_tiles, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, 8 * p, hist + 256 * p, status, tile_counter);
        *launches += 1;
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// hierarchy: thread i builds internal node i (i < n-1) and leaf i (i < n)
__global__ void k_hierarchy(const uint64_t *keys, const uint32_t *sorted_prim, const float4 *prim_lo, const float4 *prim_hi,
                            int n, BNode *bn, int *parent, int *node_count, float *node_cost, uint32_t *flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    {
        const uint32_t g = sorted_prim[i];
        const float4 a = prim_lo[g], b = prim_hi[g];
        BNode leaf;
        leaf.lo = f3(a.x, a.y, a.z); leaf.hi = f3(b.x, b.y, b.z); leaf.left = -1; leaf.right = -1;
        bn[n - 1 + i] = leaf;
        node_count[n - 1 + i] = 1;
        node_cost[n - 1 + i] = SOLB_SAH_CT * half_area(leaf.lo, leaf.hi);
    }
    if (i < n - 1) {
        int l, r, first, last;
        karras_node(keys, n, i, l, r, first, last);
        bn[i].left = l;
        bn[i].right = r;
        parent[l] = i;
        parent[r] = i;
        flags[i] = 0;
        if (i == 0) parent[0] = -1;
    }
}

// bottom-up: second thread to arrive at a node owns it.  mode 0: refit only (boxes, counts, SAH cost);
// mode 1: also restructure the treelet rooted at every node with at least `gamma` triangles.
__global__ void k_bottom_up(int n, BNode *bn, int *parent, int *node_count, float *node_cost, uint32_t *flags, int mode, int gamma) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    TreeletScratch sc;
    int node = parent[n - 1 + j];
    while (node >= 0) {
        __threadfence();
        const uint32_t old = atomicAdd(&flags[node], 1u);
        if (old == 0) return;  // the sibling subtree is not finished yet; its thread will continue
        __threadfence();
        const int l = bn[node].left, r = bn[node].right;
        const float3 lo = fmin3(bn[l].lo, bn[r].lo), hi = fmax3(bn[l].hi, bn[r].hi);
        bn[node].lo = lo;
        bn[node].hi = hi;
        const int cnt = node_count[l] + node_count[r];
        node_count[node] = cnt;
        node_cost[node] = leaf_or_internal_cost(half_area(lo, hi), node_cost[l] + node_cost[r], cnt);
        if (mode == 1 && cnt >= gamma) optimize_treelet(bn, parent, node_cost, node_count, n - 1, node, sc);
        node = parent[node];
    }
}

// ---------------------------------------------------------------------------------------------------
// Warp-cooperative treelet restructuring (Karras & Aila 2013, section 4): threads still walk up from the
// leaves with atomic flags, but when lanes of a warp reach treelet roots the warp optimises those treelets
// cooperatively (lane groups share the 2^7 subset areas and the dynamic program over subset sizes), so one
// treelet costs microseconds instead of the ~50 us of the per-thread version (k_bottom_up mode 1), which also
// shortens the serial critical path up the tree by the same factor.
struct WarpTreeletScratch {
    float area[128];
    float cost[128];
    float lo[SOLB_TREELET_N][3], hi[SOLB_TREELET_N][3];
    float lcost[SOLB_TREELET_N];
    int lcount[SOLB_TREELET_N];
    int leaves[SOLB_TREELET_N];
    int internals[SOLB_TREELET_N];
    uint8_t part[128];
    uint8_t small[128];  // subset holds <= SOLB_MAX_LEAF_TRIS triangles (count saturated)
    int nl;
    int changed;
};

// Partition tables of the 7-leaf treelet (the only size gamma >= 7 produces): the 63 splits of the full set and the 31
// splits of each of its seven 6-element subsets, in the enumeration order of the serial loop (p = (p - delta) & s), so the
// lanes of a warp can share one subset's splits and still break ties like the serial code (first in order wins).
struct TreeletTables {
    uint8_t part7[64];
    uint8_t part6[7][32];
};
__device__ void treelet_tables_init(TreeletTables &t, int tid) {
    if (tid < 8) {  // tid 0..6: the 6-element subsets, tid 7: the full set
        const int s = tid < 7 ? (127 ^ (1 << tid)) : 127;
        uint8_t *out = tid < 7 ? t.part6[tid] : t.part7;
        const int delta = (s - 1) & s;
        int p = (-delta) & s, r = 0;
        do {
            out[r++] = (uint8_t)p;
            p = (p - delta) & s;
        } while (p != 0);
    }
}

// Each warp optimises up to TL_GROUPS ready treelets at once, one per group of TL_GROUP lanes: treelet formation and the
// topology rebuild are serial chains of dependent global loads / stores run by the group's first lane, so four of them
// in flight per warp quadruple the memory-level parallelism of the pass; the 2^7 subset areas and the dynamic program are
// spread over the group's lanes.  Treelets that are ready in the same iteration root disjoint subtrees (a node only
// becomes ready after both children were finished in EARLIER iterations), so they are independent.
constexpr int TL_GROUP = 32;  // measured at 20 M triangles: 8-lane groups (4 treelets per warp) 70 ms, whole warp per treelet 48 ms
constexpr int TL_GROUPS = 32 / TL_GROUP;

// root < 0: this group has no treelet this round (it still takes part in the warp-wide barriers)
__device__ void coop_optimize_treelet(BNode *bn, int *parent, float *node_cost, int *node_count, int n_internal, int root,
                                      WarpTreeletScratch &ws, const TreeletTables &tt, int glane) {
    if (glane == 0) {  // treelet formation: expand the largest-area internal leaf until 7 leaves
        int nl = 0, ni = 0;
        if (root >= 0) {
            ws.internals[ni++] = root;
            ws.leaves[nl++] = bn[root].left;
            ws.leaves[nl++] = bn[root].right;
            while (nl < SOLB_TREELET_N) {
                int best = -1;
                float best_a = -1.0f;
                for (int i = 0; i < nl; i++) {
                    const int c = ws.leaves[i];
                    if (c < n_internal) {
                        const float a = half_area(bn[c].lo, bn[c].hi);
                        if (a > best_a) { best_a = a; best = i; }
                    }
                }
                if (best < 0) break;
                const int c = ws.leaves[best];
                ws.internals[ni++] = c;
                ws.leaves[best] = bn[c].left;
                ws.leaves[nl++] = bn[c].right;
            }
        }
        ws.nl = nl;
        ws.changed = 0;
    }
    __syncwarp();
    const int nl = ws.nl;
    const bool on = nl >= 3;
    const int full = on ? (1 << nl) - 1 : 0;
    if (on && glane < nl) {
        const BNode b = bn[ws.leaves[glane]];
        ws.lo[glane][0] = b.lo.x; ws.lo[glane][1] = b.lo.y; ws.lo[glane][2] = b.lo.z;
        ws.hi[glane][0] = b.hi.x; ws.hi[glane][1] = b.hi.y; ws.hi[glane][2] = b.hi.z;
        ws.lcost[glane] = node_cost[ws.leaves[glane]];
        ws.lcount[glane] = node_count[ws.leaves[glane]];
    }
    __syncwarp();
    for (int s = glane + 1; s <= full; s += TL_GROUP) {  // subset areas and "fits in a leaf" flags
        float lo0 = 3.4e38f, lo1 = 3.4e38f, lo2 = 3.4e38f, hi0 = -3.4e38f, hi1 = -3.4e38f, hi2 = -3.4e38f;
        int cnt = 0;
        for (int i = 0; i < nl; i++)
            if (s & (1 << i)) {
                lo0 = fminf(lo0, ws.lo[i][0]); lo1 = fminf(lo1, ws.lo[i][1]); lo2 = fminf(lo2, ws.lo[i][2]);
                hi0 = fmaxf(hi0, ws.hi[i][0]); hi1 = fmaxf(hi1, ws.hi[i][1]); hi2 = fmaxf(hi2, ws.hi[i][2]);
                cnt += min(ws.lcount[i], 64);
            }
        ws.area[s] = half_area(f3(lo0, lo1, lo2), f3(hi0, hi1, hi2));
        ws.small[s] = cnt <= SOLB_MAX_LEAF_TRIS ? (uint8_t)cnt : 0;
        if (__popc(s) == 1) { ws.cost[s] = ws.lcost[31 - __clz(s)]; ws.part[s] = 0; }
    }
    __syncwarp();
    // dynamic program over subset sizes (every group runs all the barriers).  Sizes 2..5 (21 + 35 + 35 + 21 subsets with
    // 1..15 splits each): one lane per subset.  Sizes 6 and 7 of a full 7-leaf treelet (7 subsets x 31 splits, 1 x 63) would
    // leave all but 7 / 1 lanes idle for 31 / 63 serial steps, 45 % of the kernel's instructions in the first profile
    // (profiles/r01_ncu_k_bottom_up_coop.txt): there the lanes split each subset's table of splits and reduce (cost, order) by shuffle.
    const bool fast67 = TL_GROUP == 32 && nl == SOLB_TREELET_N;  // warp-uniform only with one treelet per warp (full-mask shuffles)
    for (int k = 2; k <= SOLB_TREELET_N; k++) {
        if (fast67 && k >= 6) {
            constexpr int LPS = TL_GROUP / 8;              // lanes per 6-element subset (7 subsets, 8th share idles)
            const int lanes = k == 6 ? LPS : TL_GROUP;     // lanes sharing one subset
            const int j = k == 6 ? glane / LPS : 0, sub = k == 6 ? glane % LPS : glane;
            const int n_parts = k == 6 ? 31 : 63;
            const bool have = k == 7 || j < 7;
            const int s = k == 6 ? (127 ^ (1 << (j < 7 ? j : 0))) : 127;
            const uint8_t *table = k == 6 ? tt.part6[j < 7 ? j : 0] : tt.part7;
            float best_c = 3.4e38f;
            int best_r = 0x7fffffff;
            if (have)
                for (int r = sub; r < n_parts; r += lanes) {
                    const int p = table[r];
                    const float c = ws.cost[p] + ws.cost[s ^ p];
                    if (c < best_c) { best_c = c; best_r = r; }
                }
            for (int off = 1; off < lanes; off <<= 1) {  // lanes of one subset are contiguous and lanes is a power of two
                const float oc = __shfl_xor_sync(0xffffffffu, best_c, off, TL_GROUP);
                const int orr = __shfl_xor_sync(0xffffffffu, best_r, off, TL_GROUP);
                if (oc < best_c || (oc == best_c && orr < best_r)) { best_c = oc; best_r = orr; }
            }
            if (have && sub == 0) {
                float c = SOLB_SAH_CI * ws.area[s] + best_c;
                if (ws.small[s]) c = fminf(c, SOLB_SAH_CT * ws.area[s] * (float)ws.small[s]);
                ws.cost[s] = c;
                ws.part[s] = best_r < n_parts ? table[best_r] : (uint8_t)0;
            }
        } else if (k <= nl)
            for (int s = glane + 1; s <= full; s += TL_GROUP) {
                if (__popc(s) != k) continue;
                float best_c = 3.4e38f;
                int best_p = 0;
                const int delta = (s - 1) & s;
                int p = (-delta) & s;
                do {
                    const float c = ws.cost[p] + ws.cost[s ^ p];
                    if (c < best_c) { best_c = c; best_p = p; }
                    p = (p - delta) & s;
                } while (p != 0);
                float c = SOLB_SAH_CI * ws.area[s] + best_c;
                if (ws.small[s]) c = fminf(c, SOLB_SAH_CT * ws.area[s] * (float)ws.small[s]);
                ws.cost[s] = c;
                ws.part[s] = (uint8_t)best_p;
            }
        __syncwarp();
    }
    if (on && glane == 0 && ws.cost[full] < node_cost[root] * 0.9999f) {  // rebuild the topology, reusing the internal node ids
        int stack_set[SOLB_TREELET_N], stack_node[SOLB_TREELET_N], order[SOLB_TREELET_N];
        int sp = 0, next_internal = 1, n_order = 0;
        stack_set[sp] = full; stack_node[sp] = root; sp++;
        while (sp) {
            sp--;
            const int s = stack_set[sp], node = stack_node[sp];
            order[n_order++] = node;
            const int sub[2] = { (int)ws.part[s], s ^ (int)ws.part[s] };
            int child[2];
            for (int k = 0; k < 2; k++) {
                if (__popc(sub[k]) == 1) child[k] = ws.leaves[31 - __clz(sub[k])];
                else {
                    child[k] = ws.internals[next_internal++];
                    stack_set[sp] = sub[k]; stack_node[sp] = child[k]; sp++;
                }
                parent[child[k]] = node;
            }
            bn[node].left = child[0];
            bn[node].right = child[1];
        }
        for (int k = n_order - 1; k >= 0; k--) {
            const int node = order[k];
            const int l = bn[node].left, r = bn[node].right;
            const float3 lo = fmin3(bn[l].lo, bn[r].lo), hi = fmax3(bn[l].hi, bn[r].hi);
            bn[node].lo = lo;
            bn[node].hi = hi;
            node_count[node] = node_count[l] + node_count[r];
            node_cost[node] = leaf_or_internal_cost(half_area(lo, hi), node_cost[l] + node_cost[r], node_count[node]);
        }
        __threadfence();
    }
    __syncwarp();
}

constexpr int TL_BLOCK = 128;
__global__ void __launch_bounds__(TL_BLOCK) k_bottom_up_coop(int n, BNode *bn, int *parent, int *node_count, float *node_cost,
                                                             uint32_t *flags, int gamma) {
    __shared__ WarpTreeletScratch scratch[TL_BLOCK / 32][TL_GROUPS];
    __shared__ TreeletTables tables;
    treelet_tables_init(tables, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, group = lane / TL_GROUP, glane = lane % TL_GROUP;
    WarpTreeletScratch &ws = scratch[threadIdx.x >> 5][group];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int node = j < n ? parent[n - 1 + j] : -1;
    bool active = node >= 0;
    while (__any_sync(0xffffffffu, active)) {
        bool ready = false;
        if (active) {
            __threadfence();
            const uint32_t old = atomicAdd(&flags[node], 1u);
            if (old == 0) {
                active = false;  // the sibling subtree is not finished; its thread continues upward
            } else {
                __threadfence();
                const int l = bn[node].left, r = bn[node].right;
                const float3 lo = fmin3(bn[l].lo, bn[r].lo), hi = fmax3(bn[l].hi, bn[r].hi);
                bn[node].lo = lo;
                bn[node].hi = hi;
                const int cnt = node_count[l] + node_count[r];
                node_count[node] = cnt;
                node_cost[node] = leaf_or_internal_cost(half_area(lo, hi), node_cost[l] + node_cost[r], cnt);
                ready = cnt >= gamma;
                __threadfence();
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, active && ready);
        while (m) {  // hand the next TL_GROUPS ready treelets to the warp's lane groups
            int root = -1;
#pragma unroll
            for (int g = 0; g < TL_GROUPS; g++) {
                const int leader = m ? __ffs(m) - 1 : 0;
                const int r = __shfl_sync(0xffffffffu, node, leader);
                if (m && g == group) root = r;
                m &= m - 1;
            }
            coop_optimize_treelet(bn, parent, node_cost, node_count, n - 1, root, ws, tables, glane);
        }
        if (active) {
            node = parent[node];
            if (node < 0) active = false;
        }
    }
}

// bottom-up pass of the optimal-collapse dynamic program (DpEntry per binary node); same atomic-flag walk as k_bottom_up
__global__ void k_bottom_up_dp(int n, const BNode *bn, const int *parent, const int *node_count, uint32_t *flags, DpEntry *dp,
                               const DpCost cost) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    {
        const BNode leaf = bn[n - 1 + j];
        DpEntry e;
        dp_leaf_entry(e, half_area(leaf.lo, leaf.hi), cost);
        dp[n - 1 + j] = e;
    }
    int node = parent[n - 1 + j];
    while (node >= 0) {
        __threadfence();
        const uint32_t old = atomicAdd(&flags[node], 1u);
        if (old == 0) return;
        __threadfence();
        const BNode b = bn[node];
        const DpEntry l = dp[b.left], r = dp[b.right];
        DpEntry e;
        dp_inner_entry(e, l, r, half_area(b.lo, b.hi), node_count[node], cost);
        dp[node] = e;
        node = parent[node];
    }
}

// ---------------------------------------------------------------------------------------------------
// PLOC: parallel locally-ordered clustering (Meister & Bittner 2018; the radius and the single-pass iteration after Benthin
// et al. 2022, "PLOC++").  Replaces LBVH + refit + treelet restructuring + the bottom-up DP pass for the flattened build: the
// Morton-sorted leaves are the initial clusters; every iteration each cluster looks PLOC_R clusters to its left and right for
// the neighbour whose union with it has the smallest area, mutual nearest neighbours are merged into a new binary node, and
// the cluster array is compacted in place (order preserved).  The array shrinks by 30-45 % per iteration, ~40 iterations at
// 20 M triangles.  Box, triangle count, SAH cost and the optimal-collapse table of a new node are computed when it is
// created, so no separate bottom-up pass (a pointer chase with one atomic per node) is needed afterwards.
// The treelet pass this replaces spent 44.5 of the 62 ms of a 20 M-triangle build at 7.7 of 32 lanes active
// (profiles/r01_launch_shares_build_20m.txt, r01_ncu_k_bottom_up_coop.txt).
constexpr int PLOC_CH = 256;  // clusters per block
#ifndef SOLB_PLOC_R
#define SOLB_PLOC_R 8
#endif
constexpr int PLOC_R = SOLB_PLOC_R;  // search radius

__global__ void k_ploc_init(int n, const uint32_t *__restrict__ sorted_prim, const float4 *__restrict__ prim_lo,
                            const float4 *__restrict__ prim_hi, BNode *bn, int *parent, int *node_count, float *node_cost, DpEntry *dp,
                            const DpCost cost, float4 *c_lo, float4 *c_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = sorted_prim[i];
    const float4 a = prim_lo[g], b = prim_hi[g];
    BNode leaf;
    leaf.lo = f3(a.x, a.y, a.z); leaf.hi = f3(b.x, b.y, b.z); leaf.left = -1; leaf.right = -1;
    const int node = n - 1 + i;
    bn[node] = leaf;
    node_count[node] = 1;
    const float area = half_area(leaf.lo, leaf.hi);
    node_cost[node] = SOLB_SAH_CT * area;
    if (dp) {
        DpEntry e;
        dp_leaf_entry(e, area, cost);
        dp[node] = e;
    }
    c_lo[i] = make_float4(a.x, a.y, a.z, __int_as_float(node));
    c_hi[i] = make_float4(b.x, b.y, b.z, __int_as_float(1));
    if (i == 0) parent[0] = -1;
}

// One PLOC iteration over n_clusters clusters: nearest neighbours, merges, block-local compaction into tmp (block b's
// survivors at tmp[b * PLOC_CH ...]) and their number in block_counts[b].
__global__ void __launch_bounds__(PLOC_CH) k_ploc_merge(int n_prims, int n_clusters, const float4 *__restrict__ in_lo,
                                                        const float4 *__restrict__ in_hi, float4 *tmp_lo, float4 *tmp_hi,
                                                        uint32_t *block_counts, uint32_t *node_counter, BNode *bn, int *parent,
                                                        int *node_count, float *node_cost, DpEntry *dp, const DpCost cost) {
    constexpr int W = PLOC_CH + 4 * PLOC_R;  // the chunk and 2 R clusters on either side
    __shared__ float4 s_lo[W], s_hi[W];
    __shared__ int s_nn[PLOC_CH + 2 * PLOC_R];  // nearest neighbour (global index) of local items R .. R + CH + 2 R - 1
    __shared__ uint32_t s_warp_valid[PLOC_CH / 32], s_warp_merge[PLOC_CH / 32], s_node_base;
    const int tid = threadIdx.x;
    const int g0 = blockIdx.x * PLOC_CH - 2 * PLOC_R;  // global index of local item 0
    for (int l = tid; l < W; l += PLOC_CH) {
        const int g = g0 + l;
        if (g >= 0 && g < n_clusters) { s_lo[l] = in_lo[g]; s_hi[l] = in_hi[g]; }
    }
    __syncthreads();
    // nearest neighbour of the chunk's clusters and of the R clusters on either side (their choice decides mutuality).
    // Ties go to the smaller index, which is the symmetric order (area, min(i, j), max(i, j)): the globally smallest pair is
    // always mutual, so every iteration merges at least once.
    for (int q = tid; q < PLOC_CH + 2 * PLOC_R; q += PLOC_CH) {
        const int l = PLOC_R + q, g = g0 + l;
        int best = -1;
        if (g >= 0 && g < n_clusters) {
            const float4 lo = s_lo[l], hi = s_hi[l];
            float best_a = 3.4e38f;
#pragma unroll
            for (int d = -PLOC_R; d <= PLOC_R; d++) {
                if (d == 0) continue;
                const int gj = g + d;
                if (gj < 0 || gj >= n_clusters) continue;
                const float4 lo2 = s_lo[l + d], hi2 = s_hi[l + d];
                const float ex = fmaxf(hi.x, hi2.x) - fminf(lo.x, lo2.x), ey = fmaxf(hi.y, hi2.y) - fminf(lo.y, lo2.y),
                            ez = fmaxf(hi.z, hi2.z) - fminf(lo.z, lo2.z);
                const float a = ex * ey + ey * ez + ez * ex;
                if (a < best_a) { best_a = a; best = gj; }
            }
        }
        s_nn[q] = best;
    }
    __syncthreads();
    const int l = 2 * PLOC_R + tid, g = g0 + l;  // this thread's cluster
    const bool have = g < n_clusters;
    int partner = -1;
    if (have) {
        const int j = s_nn[PLOC_R + tid];
        if (j >= 0 && s_nn[j - g0 - PLOC_R] == g) partner = j;
    }
    const bool winner = partner > g;             // the left cluster of a mutual pair creates the node
    const bool valid = have && !(partner >= 0 && partner < g);  // the right one disappears
    // ranks inside the block
    const uint32_t lane = tid & 31, wid = tid >> 5;
    const uint32_t b_valid = __ballot_sync(0xffffffffu, valid), b_merge = __ballot_sync(0xffffffffu, winner);
    if (lane == 0) { s_warp_valid[wid] = __popc(b_valid); s_warp_merge[wid] = __popc(b_merge); }
    __syncthreads();
    uint32_t valid_before = 0, merge_before = 0, valid_total = 0, merge_total = 0;
#pragma unroll
    for (int w = 0; w < PLOC_CH / 32; w++) {
        if (w < (int)wid) { valid_before += s_warp_valid[w]; merge_before += s_warp_merge[w]; }
        valid_total += s_warp_valid[w];
        merge_total += s_warp_merge[w];
    }
    if (tid == 0) {
        s_node_base = merge_total ? atomicAdd(node_counter, merge_total) : 0u;
        block_counts[blockIdx.x] = valid_total;
    }
    __syncthreads();
    if (!valid) return;
    float4 lo = s_lo[l], hi = s_hi[l];
    if (winner) {
        const float4 lo2 = s_lo[partner - g0], hi2 = s_hi[partner - g0];
        const int left = __float_as_int(lo.w), right = __float_as_int(lo2.w);
        const int cnt = __float_as_int(hi.w) + __float_as_int(hi2.w);
        // internal nodes are numbered from the top down so that the last merge (the root) gets index 0 (Karras numbering)
        const int node = (n_prims - 2) - (int)(s_node_base + merge_before + __popc(b_merge & ((1u << lane) - 1u)));
        BNode b;
        b.lo = f3(fminf(lo.x, lo2.x), fminf(lo.y, lo2.y), fminf(lo.z, lo2.z));
        b.hi = f3(fmaxf(hi.x, hi2.x), fmaxf(hi.y, hi2.y), fmaxf(hi.z, hi2.z));
        b.left = left; b.right = right;
        bn[node] = b;
        parent[left] = node;
        parent[right] = node;
        node_count[node] = cnt;
        const float area = half_area(b.lo, b.hi);
        node_cost[node] = leaf_or_internal_cost(area, node_cost[left] + node_cost[right], cnt);
        if (dp) {
            const DpEntry dl = dp[left], dr = dp[right];
            DpEntry e;
            dp_inner_entry(e, dl, dr, area, cnt, cost);
            dp[node] = e;
        }
        lo = make_float4(b.lo.x, b.lo.y, b.lo.z, __int_as_float(node));
        hi = make_float4(b.hi.x, b.hi.y, b.hi.z, __int_as_float(cnt));
    }
    const uint32_t rank = valid_before + __popc(b_valid & ((1u << lane) - 1u));
    tmp_lo[(size_t)blockIdx.x * PLOC_CH + rank] = lo;
    tmp_hi[(size_t)blockIdx.x * PLOC_CH + rank] = hi;
}

// exclusive scan of the per-block survivor counts (one block); total -> *n_out
__global__ void __launch_bounds__(1024) k_ploc_scan(const uint32_t *__restrict__ block_counts, uint32_t n_blocks, uint32_t *offsets,
                                                    uint32_t *n_out) {
    __shared__ uint32_t s_part[1024];
    const uint32_t tid = threadIdx.x, per = (n_blocks + 1023u) / 1024u;
    const uint32_t b0 = tid * per, b1 = min(b0 + per, n_blocks);
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b1; b++) sum += block_counts[b];
    s_part[tid] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024u; d <<= 1) {  // Hillis-Steele inclusive scan
        const uint32_t v = tid >= d ? s_part[tid - d] : 0u;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint32_t run = s_part[tid] - sum;
    for (uint32_t b = b0; b < b1; b++) { offsets[b] = run; run += block_counts[b]; }
    if (tid == 1023u) *n_out = s_part[1023];
}

// block b's survivors -> their final, order-preserving position
__global__ void __launch_bounds__(PLOC_CH) k_ploc_scatter(const float4 *__restrict__ tmp_lo, const float4 *__restrict__ tmp_hi,
                                                          const uint32_t *__restrict__ block_counts, const uint32_t *__restrict__ offsets,
                                                          float4 *out_lo, float4 *out_hi) {
    const uint32_t t = threadIdx.x;
    if (t >= block_counts[blockIdx.x]) return;
    const size_t src = (size_t)blockIdx.x * PLOC_CH + t, dst = (size_t)offsets[blockIdx.x] + t;
    out_lo[dst] = tmp_lo[src];
    out_hi[dst] = tmp_hi[src];
}

__global__ void k_clear_u32(uint32_t *p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

__global__ void k_collapse_level(const BNode *bn, const int *node_count, int n_internal, const CollapseItem *queue_in,
                                 uint32_t n_items, Node8 *wide, uint32_t *counters /*0: wide, 1: tris, 2: queue_out*/,
                                 const uint32_t *sorted_prim, const Tri48 *tri_world, Tri48 *tri_out, CollapseItem *queue_out,
                                 uint32_t *leaf_prim_out, const DpEntry *dp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    collapse_one(bn, node_count, n_internal, queue_in[i], wide, &counters[0], &counters[1], sorted_prim, tri_world, tri_out,
                 queue_out, &counters[2], leaf_prim_out, dp);
}

// n == 1: a root with one single-triangle leaf
__global__ void k_single_tri_root(const float4 *prim_lo, const float4 *prim_hi, const Tri48 *tri_world, Node8 *wide, Tri48 *tri_out) {
    ChildRef ch[8];
    for (int s = 0; s < 8; s++) ch[s].valid = 0;
    const float3 lo = f3(prim_lo[0].x, prim_lo[0].y, prim_lo[0].z), hi = f3(prim_hi[0].x, prim_hi[0].y, prim_hi[0].z);
    ch[0].valid = 1; ch[0].lo = lo; ch[0].hi = hi; ch[0].is_inner = 0; ch[0].tri_offset = 0; ch[0].tri_count = 1;
    encode_node8(wide[0], lo, hi, 0, 0, ch);
    tri_out[0] = tri_world[0];
}

__global__ void k_empty_root(Node8 *wide) {
    ChildRef ch[8];
    for (int s = 0; s < 8; s++) ch[s].valid = 0;
    encode_node8(wide[0], f3(0, 0, 0), f3(0, 0, 0), 0, 0, ch);
}

// ---------------------------------------------------------------------------------------------------
// two-level build kernels (SOLB_ACCEL_TWO_LEVEL)

// bounds[6 * b + 0..2] = min centroid (ordered), [3..5] = max centroid (ordered)
__global__ void k_init_bounds_n(uint32_t *bounds, uint32_t n_sets) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 6 * n_sets) bounds[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

// One thread per GEOMETRY triangle g (all BLASes back to back): object-space record, box, per-BLAS centroid bounds.
__global__ void k_prep_tris_obj(const DeviceSceneView sv, Tri48 *tri_obj, float4 *prim_lo, float4 *prim_hi, uint32_t *blas_bounds) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < sv.n_geom_tris;
    uint32_t b = 0xffffffffu;
    float3 c = f3(0, 0, 0);
    if (valid) {
        uint32_t lo = 0, hi = sv.n_blas;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sv.blas[mid].first_tri <= g) lo = mid; else hi = mid;
        }
        b = lo;
        const DeviceBlas db = sv.blas[b];
        const uint32_t prim = g - db.first_tri;
        float3 p[3];
        for (int k = 0; k < 3; k++) {
            const uint32_t vi = db.first_vertex + sv.indices[db.first_index + 3 * prim + k];
            const float4 pos = sv.vertices[4 * (size_t)vi];
            p[k] = f3(pos.x, pos.y, pos.z);
        }
        Tri48 t;
        t.v0 = make_float4(p[0].x, p[0].y, p[0].z, __uint_as_float(b));  // instance id comes from the TLAS leaf
        t.v1 = make_float4(p[1].x, p[1].y, p[1].z, __uint_as_float(prim));
        t.v2 = make_float4(p[2].x, p[2].y, p[2].z, __uint_as_float(g));
        tri_obj[g] = t;
        const float3 lo3 = fmin3(p[0], fmin3(p[1], p[2])), hi3 = fmax3(p[0], fmax3(p[1], p[2]));
        prim_lo[g] = make_float4(lo3.x, lo3.y, lo3.z, __uint_as_float(b));
        prim_hi[g] = make_float4(hi3.x, hi3.y, hi3.z, 0.0f);
        c = (lo3 + hi3) * 0.5f;
    }
    // warps that sit inside one BLAS reduce first (the common case: BLASes hold thousands of triangles)
    const uint32_t b0 = __shfl_sync(0xffffffffu, b, 0);
    if (__all_sync(0xffffffffu, b == b0) && b0 != 0xffffffffu) {
        float mn[3] = { c.x, c.y, c.z }, mx[3] = { c.x, c.y, c.z };
        for (int off = 16; off; off >>= 1)
            for (int k = 0; k < 3; k++) {
                mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], off));
                mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
            }
        if ((threadIdx.x & 31) == 0)
            for (int k = 0; k < 3; k++) {
                atomicMin(&blas_bounds[6 * b0 + k], float_to_ordered(mn[k]));
                atomicMax(&blas_bounds[6 * b0 + 3 + k], float_to_ordered(mx[k]));
            }
    } else if (valid) {
        const float cc[3] = { c.x, c.y, c.z };
        for (int k = 0; k < 3; k++) {
            atomicMin(&blas_bounds[6 * b + k], float_to_ordered(cc[k]));
            atomicMax(&blas_bounds[6 * b + 3 + k], float_to_ordered(cc[k]));
        }
    }
}

// Segmented Morton key: [BLAS id | Morton code of the centroid inside its BLAS's centroid box], so one sort and one
// Karras hierarchy over all triangles yield every BLAS as its own subtree (keys sharing a prefix form a subtree).
__global__ void k_morton_seg(const float4 *prim_lo, const float4 *prim_hi, uint32_t n, const uint32_t *blas_bounds, int morton_bits,
                             uint64_t *keys, uint32_t *vals) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const float4 a = prim_lo[g], b4 = prim_hi[g];
    const uint32_t b = __float_as_uint(a.w);
    const uint32_t *bb = blas_bounds + 6 * (size_t)b;
    const float3 lo = f3(ordered_to_float(bb[0]), ordered_to_float(bb[1]), ordered_to_float(bb[2]));
    const float3 hi = f3(ordered_to_float(bb[3]), ordered_to_float(bb[4]), ordered_to_float(bb[5]));
    const float3 ext = hi - lo;
    const float3 inv = f3(ext.x > 0.0f ? 1.0f / ext.x : 0.0f, ext.y > 0.0f ? 1.0f / ext.y : 0.0f, ext.z > 0.0f ? 1.0f / ext.z : 0.0f);
    const float3 c = f3((a.x + b4.x) * 0.5f, (a.y + b4.y) * 0.5f, (a.z + b4.z) * 0.5f);
    keys[g] = ((uint64_t)b << morton_bits) | (morton63(c, lo, inv) >> (63 - morton_bits));
    vals[g] = g;
}

// BLAS b covers sorted positions [first, last]; its subtree root is the internal node whose Karras range is exactly
// that (node `first` or node `last`), or the leaf itself for a single-triangle BLAS.  Detach it from the levels above.
__global__ void k_blas_roots(const uint64_t *keys, int n, const DeviceBlas *blas, uint32_t n_blas, int *parent, int *blas_root) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blas) return;
    const int first = (int)blas[b].first_tri, cnt = (int)(blas[b].n_indices / 3), last = first + cnt - 1;
    int root;
    if (cnt == 1) root = n - 1 + first;
    else {
        root = -1;
        for (int k = 0; k < 2 && root < 0; k++) {
            const int i = k ? last : first;
            if (i > n - 2) continue;
            int l, r, f, e;
            karras_node(keys, n, i, l, r, f, e);
            if (f == first && e == last) root = i;
        }
    }
    blas_root[b] = root;
    if (root >= 0) parent[root] = -1;
}

__global__ void k_blas_boxes(const BNode *bn, const int *blas_root, uint32_t n_blas, float4 *blas_box) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blas) return;
    const BNode r = bn[blas_root[b]];
    blas_box[2 * b] = make_float4(r.lo.x, r.lo.y, r.lo.z, 0.0f);
    blas_box[2 * b + 1] = make_float4(r.hi.x, r.hi.y, r.hi.z, 0.0f);
}

// collapse seeds: item i turns the root of multi-triangle BLAS list[i] into wide node tlas_cap + list[i]
__global__ void k_seed_blas_collapse(const uint32_t *list, uint32_t n_items, const int *blas_root, uint32_t tlas_cap, CollapseItem *queue) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    CollapseItem it;
    it.bnode = blas_root[list[i]];
    it.wnode = tlas_cap + list[i];
    queue[i] = it;
}

// single-triangle BLAS list[i]: a root with one single-triangle leaf, triangle slot i
__global__ void k_single_tri_blas(const uint32_t *list, uint32_t n_items, const DeviceBlas *blas, const float4 *prim_lo,
                                  const float4 *prim_hi, const Tri48 *tri_obj, uint32_t tlas_cap, Node8 *wide, Tri48 *tri_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const uint32_t b = list[i], g = blas[b].first_tri;
    ChildRef ch[8];
    for (int s = 0; s < 8; s++) ch[s].valid = 0;
    const float3 lo = f3(prim_lo[g].x, prim_lo[g].y, prim_lo[g].z), hi = f3(prim_hi[g].x, prim_hi[g].y, prim_hi[g].z);
    ch[0].valid = 1; ch[0].lo = lo; ch[0].hi = hi; ch[0].is_inner = 0; ch[0].tri_offset = 0; ch[0].tri_count = 1;
    encode_node8(wide[tlas_cap + b], lo, hi, 0, i, ch);
    tri_out[i] = tri_obj[g];
}

// TLAS primitives: world box of every instance (its BLAS's root box under the instance matrix) + centroid bounds
__global__ void k_inst_boxes(const DeviceInstance *instances, uint32_t n_inst, const float4 *blas_box, float4 *prim_lo, float4 *prim_hi,
                             uint32_t *bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst) return;
    const DeviceInstance &di = instances[i];
    const float4 a = blas_box[2 * di.blas], b = blas_box[2 * di.blas + 1];
    float3 lo, hi;
    transform_box(di.transform, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), lo, hi);
    prim_lo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
    prim_hi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    const float cc[3] = { (lo.x + hi.x) * 0.5f, (lo.y + hi.y) * 0.5f, (lo.z + hi.z) * 0.5f };
    for (int k = 0; k < 3; k++) {
        atomicMin(&bounds[k], float_to_ordered(cc[k]));
        atomicMax(&bounds[3 + k], float_to_ordered(cc[k]));
    }
}

// TLAS leaf slot j holds instance leaf_prim[j]
__global__ void k_write_inst_leaves(const uint32_t *leaf_prim, uint32_t n_inst, const DeviceInstance *instances, uint32_t tlas_cap,
                                    InstLeaf *out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_inst) return;
    const uint32_t i = leaf_prim ? leaf_prim[j] : j;
    out[j] = make_inst_leaf(instances[i].transform, tlas_cap + instances[i].blas, i);
}

// n_inst == 1: TLAS root with one single-instance leaf
__global__ void k_single_inst_root(const float4 *prim_lo, const float4 *prim_hi, Node8 *wide) {
    ChildRef ch[8];
    for (int s = 0; s < 8; s++) ch[s].valid = 0;
    const float3 lo = f3(prim_lo[0].x, prim_lo[0].y, prim_lo[0].z), hi = f3(prim_hi[0].x, prim_hi[0].y, prim_hi[0].z);
    ch[0].valid = 1; ch[0].lo = lo; ch[0].hi = hi; ch[0].is_inner = 0; ch[0].tri_offset = 0; ch[0].tri_count = 1;
    encode_node8(wide[0], lo, hi, 0, 0, ch);
}

// ---------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------
// collapse_one for a whole WARP (single-CTA TLAS kernel, optimal-collapse table present).  One thread running collapse_one
// takes ~45 000 cycles (the slot auction is 8 rounds over 64 pairs, the encoder loops over children x axes with
// divisions), and the level-synchronous collapse of a 1 000-instance TLAS waited for that five times: 64 % of the kernel
// (SOLB_TLAS_TRACE).  Here the 32 lanes split the work and write bit-identical nodes:
//   * children: the DP's expansion runs level by level (every candidate with budget left splits in the same round);
//   * slot auction: lane = (child i, slot pair), two scores per lane and a warp arg-max per round, ties to the smaller
//     (child, slot) like the serial loops;
//   * encoding: lane = (slot, axis): its own exponent requirement (max over slots by shuffle = the serial bump loop), its
//     two quantised planes with the serial conservative fix-ups, written as bytes straight into the node.
// Must be called by all 32 lanes with the same item.  dp may live in shared or global memory.
__device__ __forceinline__ void collapse_one_warp(const BNode *bn, const int *node_count, int n_internal, CollapseItem item, Node8 *wide,
                                                  uint32_t *wide_count, uint32_t *tri_count, const uint32_t *sorted_prim,
                                                  CollapseItem *queue_out, uint32_t *queue_out_count, uint32_t *leaf_prim_out,
                                                  const DpEntry *dp, int2 *scratch /* 8 entries of shared memory owned by this warp */) {
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const unsigned FULL = 0xffffffffu;
    // ---- 1. children of the wide node: lanes 0..7 hold (node, budget) in left-to-right order ----
    int c_node = -1, c_budget = 0;
    {
        const BNode self = bn[item.bnode];
        const int k = (int)((dp[item.bnode].k >> 18) & 7u);  // distribute(root, 8)
        if (lane == 0) { c_node = self.left; c_budget = k; }
        if (lane == 1) { c_node = self.right; c_budget = 8 - k; }
    }
    for (;;) {
        // a candidate splits when it is an internal node that the table gives more than one child slot
        int split_k = 0, b = c_budget > 7 ? 7 : c_budget, l = -1, r = -1;
        if (c_node >= 0 && c_node < n_internal && b > 1) {
            const uint32_t kb = dp[c_node].k;
            while (b > 1 && ((kb >> (3 * (b - 2))) & 7u) == 0u) b--;
            if (b > 1) { split_k = (int)((kb >> (3 * (b - 2))) & 7u); l = bn[c_node].left; r = bn[c_node].right; }
        }
        const uint32_t m_split = __ballot_sync(FULL, split_k != 0);
        if (m_split == 0u) break;
        // new position: every splitting candidate to the left shifts this one right by one.  The candidates (at most 8, on
        // lanes 0..7) write themselves - or their two halves - to the warp's scratch row and every lane reads its new one back:
        // two stores and a load per round where a shuffle scatter took 56 shuffles (the expansion was 40 % of a call).
        const uint32_t m_have = __ballot_sync(FULL, c_node >= 0);
        const int pos = (int)lane + __popc(m_split & lt_mask);
        if (c_node >= 0) {
            if (split_k) { scratch[pos] = make_int2(l, split_k); scratch[pos + 1] = make_int2(r, b - split_k); }
            else scratch[pos] = make_int2(c_node, c_budget);
        }
        __syncwarp();
        const int n_now = __popc(m_have) + __popc(m_split);
        const int2 mine = (int)lane < n_now ? scratch[lane] : make_int2(-1, 0);
        c_node = mine.x; c_budget = mine.y;
        __syncwarp();
    }
    const int n = __popc(__ballot_sync(FULL, c_node >= 0));  // 2..8 children, on lanes 0..n-1
    // ---- 2. per child (lane i < n): box, count, leaf flag ----
    const BNode self = bn[item.bnode];
    const float3 nlo = self.lo, nhi = self.hi;
    const float3 nc = (nlo + nhi) * 0.5f;
    float3 clo = f3(0, 0, 0), chi = f3(0, 0, 0), rel = f3(0, 0, 0);
    int ccount = 0;
    bool cleaf = false;
    if ((int)lane < n) {
        const BNode b = bn[c_node];
        clo = b.lo; chi = b.hi;
        rel = (b.lo + b.hi) * 0.5f - nc;
        ccount = node_count[c_node];
        cleaf = (dp[c_node].k >> 31) != 0u;
    }
    // ---- 3. slot auction: lane = (child i = lane >> 2, slots 2 (lane & 3), 2 (lane & 3) + 1) ----
    int my_slot = -1;  // lane i < n: the slot of child i
    {
        const int ai = (int)(lane >> 2), s0 = 2 * (int)(lane & 3u);
        const float rx = __shfl_sync(FULL, rel.x, ai), ry = __shfl_sync(FULL, rel.y, ai), rz = __shfl_sync(FULL, rel.z, ai);
        float v0, v1;
        {
            const int sl = s0;
            v0 = ((sl & 4) ? rx : -rx) + ((sl & 2) ? ry : -ry) + ((sl & 1) ? rz : -rz);
        }
        {
            const int sl = s0 + 1;
            v1 = ((sl & 4) ? rx : -rx) + ((sl & 2) ? ry : -ry) + ((sl & 1) ? rz : -rz);
        }
        uint32_t free_child = (1u << n) - 1u, free_slot = 0xffu;
        for (int round = 0; round < n; round++) {
            float best_v = -3.4e38f;
            int best_idx = 64;  // i * 8 + slot; 64 = none
            if ((free_child >> ai) & 1u) {
                if (((free_slot >> s0) & 1u) && v0 > best_v) { best_v = v0; best_idx = ai * 8 + s0; }
                if (((free_slot >> (s0 + 1)) & 1u) && v1 > best_v) { best_v = v1; best_idx = ai * 8 + s0 + 1; }
            }
            // warp arg-max: the largest score, ties to the smaller (child, slot) - two warp reductions (REDUX) on an
            // order-preserving integer image of the score (+ 0.0f: -0 and +0 are one score) instead of five shuffle levels:
            // the collapse of a 1 000-instance TLAS went from 150 000 to 115 000 cycles (profiles/r02_tlas_regen.txt)
            {
                const uint32_t key = best_idx < 64 ? float_to_ordered(best_v + 0.0f) : 0u;
                const uint32_t kmax = __reduce_max_sync(FULL, key);
                best_idx = (int)__reduce_min_sync(FULL, (key == kmax && best_idx < 64) ? (uint32_t)best_idx : 64u);
            }
            const int bi = best_idx >> 3, bs = best_idx & 7;
            if ((int)lane == bi) my_slot = bs;
            free_child &= ~(1u << bi);
            free_slot &= ~(1u << bs);
        }
    }
    // ---- 4. per slot (leader lane = 4 * slot): which child, leaf / inner, triangle offsets in slot order ----
    const int sl = (int)(lane >> 2), axis = (int)(lane & 3u);
    int child = -1;  // child index held by slot sl
    for (int i = 0; i < 8; i++) {
        const int s_i = __shfl_sync(FULL, my_slot, i);
        if (i < n && s_i == sl) child = i;
    }
    const int src = child >= 0 ? child : 0;
    const float lo_x = __shfl_sync(FULL, clo.x, src), lo_y = __shfl_sync(FULL, clo.y, src), lo_z = __shfl_sync(FULL, clo.z, src);
    const float hi_x = __shfl_sync(FULL, chi.x, src), hi_y = __shfl_sync(FULL, chi.y, src), hi_z = __shfl_sync(FULL, chi.z, src);
    const int s_count = __shfl_sync(FULL, ccount, src), s_node = __shfl_sync(FULL, c_node, src);
    const bool s_leaf = __shfl_sync(FULL, (int)cleaf, src) != 0;
    const bool valid = child >= 0, inner = valid && !s_leaf, leaf = valid && s_leaf;
    const uint32_t m_inner = __ballot_sync(FULL, inner && axis == 0), m_leader = 0x11111111u;
    uint32_t tri_offset = 0, n_tris = 0, count0 = 0;
    for (int s2 = 0; s2 < 8; s2++) {
        const int cnt = __shfl_sync(FULL, leaf ? s_count : 0, 4 * s2);
        if (s2 < sl) tri_offset += (uint32_t)cnt;
        if (s2 < 4) count0 += (uint32_t)cnt;
        n_tris += (uint32_t)cnt;
    }
    const uint32_t n_inner = (uint32_t)__popc(m_inner);
    uint32_t child_base = 0, tri_base = 0, q = 0;
    if (lane == 0) {
        child_base = n_inner ? atomicAdd(wide_count, n_inner) : 0u;
        tri_base = n_tris ? atomicAdd(tri_count, n_tris) : 0u;
        q = n_inner ? atomicAdd(queue_out_count, n_inner) : 0u;
    }
    child_base = __shfl_sync(FULL, child_base, 0);
    tri_base = __shfl_sync(FULL, tri_base, 0);
    q = __shfl_sync(FULL, q, 0);
    // ---- 5. encode (bvh.cuh encode_node8, same arithmetic): lane = (slot, axis) ----
    uint8_t *node_bytes = (uint8_t *)&wide[item.wnode];
    uint32_t e_axis = 1u;
    if (axis < 3) {
        const float l = axis == 0 ? nlo.x : (axis == 1 ? nlo.y : nlo.z), h = axis == 0 ? nhi.x : (axis == 1 ? nhi.y : nhi.z);
        const float ch = axis == 0 ? hi_x : (axis == 1 ? hi_y : hi_z);
        uint32_t e = quant_exponent(h - l);
        if (valid) {  // the serial loop bumps the exponent until every child fits: the maximum of the per-child requirements
            for (;;) {
                const float inv = 1.0f / u2f(e << 23);
                if (!(ceilf((ch - l) * inv) > 255.0f) || e >= 254u) break;
                e++;
            }
        }
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) e = max(e, __shfl_xor_sync(0x77777777u, e, off));
        e_axis = e;
        uint32_t ql_u = 0, qh_u = 0;
        if (valid) {
            const float cl = axis == 0 ? lo_x : (axis == 1 ? lo_y : lo_z);
            const float sc = u2f(e << 23);
            float ql = floorf((cl - l) / sc), qh = ceilf((ch - l) / sc);
            ql = fminf(fmaxf(ql, 0.0f), 255.0f);
            qh = fminf(fmaxf(qh, 0.0f), 255.0f);
            while (ql > 0.0f && l + ql * sc > cl) ql -= 1.0f;
            while (qh < 255.0f && l + qh * sc < ch) qh += 1.0f;
            ql_u = (uint32_t)ql; qh_u = (uint32_t)qh;
        }
        node_bytes[32 + 8 * axis + sl] = (uint8_t)ql_u;
        node_bytes[56 + 8 * axis + sl] = (uint8_t)qh_u;
    } else {
        uint32_t meta = 0;  // tmask byte of the slot
        if (leaf) meta = (s_count == 1 ? 1u : 3u) << (tri_offset - (sl < 4 ? 0u : count0));
        node_bytes[24 + sl] = (uint8_t)meta;
    }
    {
        const uint32_t ex = __shfl_sync(FULL, e_axis, 0), ey = __shfl_sync(FULL, e_axis, 1), ez = __shfl_sync(FULL, e_axis, 2);
        uint32_t imask = 0;
        for (int s2 = 0; s2 < 8; s2++) imask |= ((m_inner >> (4 * s2)) & 1u) << s2;
        if (lane == 0) {
            uint32_t *w = (uint32_t *)node_bytes;
            w[0] = f2u(nlo.x); w[1] = f2u(nlo.y); w[2] = f2u(nlo.z);
            w[3] = ex | (ey << 8) | (ez << 16) | (imask << 24);
            w[4] = child_base;
            w[5] = (tri_base & SOLB_TRI_BASE_MASK) | (count0 << 28);
        }
    }
    // ---- 6. next level's work items / leaf records (slot leaders) ----
    if (axis == 0 && valid) {
        if (inner) {
            CollapseItem it;
            it.bnode = s_node;
            it.wnode = child_base + (uint32_t)__popc(m_inner & m_leader & lt_mask);
            queue_out[q + (uint32_t)__popc(m_inner & m_leader & lt_mask)] = it;
        } else {
            int prims[SOLB_MAX_LEAF_TRIS + 1];
            const int m = gather_leaf_tris(bn, n_internal, s_node, prims);
            for (int j = 0; j < m; j++) leaf_prim_out[tri_base + tri_offset + (uint32_t)j] = sorted_prim[prims[j]];
        }
    }
}

// blas id + transform of instance i through 8-byte loads: the 200-byte records are 8-byte aligned, and 17 scalar loads at a
// 200-byte stride cost the single SM of the TLAS kernel one L1 transaction per lane each (the box and leaf-record phases were
// bound by that)
__device__ __forceinline__ uint32_t load_instance_placement(const DeviceInstance *instances, uint32_t i, float t[16]) {
    static_assert(sizeof(DeviceInstance) % 8 == 0 && offsetof(DeviceInstance, blas) == 16 && offsetof(DeviceInstance, transform) == 24,
                  "DeviceInstance layout");
    const uint2 *p = (const uint2 *)((const char *)(instances + i) + 16);
    const uint2 head = __ldg(p);  // blas, shade_first_tri
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint2 v = __ldg(p + 1 + k);
        t[2 * k] = __uint_as_float(v.x);
        t[2 * k + 1] = __uint_as_float(v.y);
    }
    return head.x;
}

// Single-CTA TLAS build for up to TLAS_FAST_MAX instances: TLAS::regenerate runs every frame in the reference
// (examples/5-pathtrace.rs:316), so for the instance counts it is used with (8 ... a few thousand) the whole rebuild is ONE
// launch working out of shared memory: instance boxes -> Morton keys -> bitonic sort -> Karras links -> atomic-flag refit ->
// level-synchronous 8-wide collapse -> instance leaf records.  No treelet pass (a plain LBVH over instance boxes).
constexpr uint32_t TLAS_FAST_MAX = 2048;
constexpr int TLAS_FAST_THREADS = 256;        // serial collapse_one (greedy collapse): few, fat threads, it wants 230 registers
constexpr int TLAS_FAST_THREADS_COOP = 1024;  // warp-cooperative collapse (optimal-collapse table present): 64 registers

struct TlasFastInfo {
    uint32_t n_wide, depth, leaf_count, pad;
    float lo[4], hi[4];
    long long t[16];  // clock64() of thread 0 at the phase boundaries (SOLB_TLAS_TRACE=1 prints the shares)
};

__host__ __device__ inline size_t tlas_fast_smem_bytes(uint32_t n) {
    uint32_t p2 = 2;
    while (p2 < n) p2 <<= 1;
    // bn[2n-1] 32 B | keys[p2] 8 B | idx[p2] 4 B | count[2n-1] 4 B | flags[n] 4 B | queues 2 x (n/2+1) x 8 B | leaf_prim[n] 4 B | parent[2n-1] 2 B
    return (size_t)(2 * n) * 32 + (size_t)p2 * 12 + (size_t)(2 * n) * 4 + (size_t)n * 4 + (size_t)(n / 2 + 2) * 16 + (size_t)n * 4 +
           (size_t)(2 * n) * 2 + 64;
}
// ... plus the optimal-collapse table (32 B per binary node) when it fits next to the rest: walking it bottom-up from global
// memory cost 49 000 of the kernel's 430 000 cycles at 1 000 instances
__host__ __device__ inline size_t tlas_fast_dp_offset(uint32_t n) { return (tlas_fast_smem_bytes(n) + 31) & ~(size_t)31; }
__host__ __device__ inline size_t tlas_fast_smem_bytes_dp(uint32_t n) { return tlas_fast_dp_offset(n) + (size_t)(2 * n) * sizeof(DpEntry); }
constexpr size_t TLAS_FAST_SMEM_LIMIT = 227 * 1024 - 512 - 2048;  // dynamic + the kernel's static shared memory (incl. 2 KB of warp scratch rows) must fit 227 KB

template <bool COOP>
__global__ void __launch_bounds__(COOP ? TLAS_FAST_THREADS_COOP : TLAS_FAST_THREADS)
k_tlas_build_small(const DeviceInstance *__restrict__ instances, uint32_t n, const float4 *__restrict__ blas_box, uint32_t tlas_cap,
                   Node8 *wide, InstLeaf *inst_leaves, TlasFastInfo *info,
                   DpEntry *dp /* [2 n - 1] global scratch; null: greedy collapse (COOP = false) or table in shared memory */,
                   const int dp_in_smem, const DpCost cost) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_bounds[6];
    __shared__ uint32_t s_counters[3];  // wide nodes, leaf slots, next queue size
    __shared__ uint32_t s_items, s_depth;
    __shared__ int2 s_expand[COOP ? TLAS_FAST_THREADS_COOP / 32 : 1][8];  // collapse_one_warp: a scratch row per warp
    uint32_t p2 = 2;
    while (p2 < n) p2 <<= 1;
    BNode *bn = (BNode *)smem;
    uint64_t *keys = (uint64_t *)(bn + 2 * n);
    uint32_t *idx = (uint32_t *)(keys + p2);
    int *count = (int *)(idx + p2);
    uint32_t *flags = (uint32_t *)(count + 2 * n);
    CollapseItem *queue_a = (CollapseItem *)(flags + n);
    CollapseItem *queue_b = queue_a + (n / 2 + 2);
    uint32_t *leaf_prim = (uint32_t *)(queue_b + (n / 2 + 2));
    uint16_t *parent = (uint16_t *)(leaf_prim + n);
    if (dp_in_smem) dp = (DpEntry *)(smem + tlas_fast_dp_offset(n));
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const int ni = (int)n - 1;
    int n_stamp = 0;
#define TLAS_STAMP() do { if (tid == 0 && n_stamp < 16) info->t[n_stamp] = clock64(); n_stamp++; } while (0)
    TLAS_STAMP();  // 0: start

    if (tid < 3) s_bounds[tid] = 0xffffffffu;
    else if (tid < 6) s_bounds[tid] = 0u;
    __syncthreads();
    // 1. instance boxes: kept in the (not yet linked) leaf slots in INSTANCE order for the moment
    {
        uint32_t bmin[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu }, bmax[3] = { 0u, 0u, 0u };
        for (uint32_t i = tid; i < n; i += nt) {
            float xf[16];
            const uint32_t blas = load_instance_placement(instances, i, xf);
            const float4 a = blas_box[2 * blas], b = blas_box[2 * blas + 1];
            float3 lo, hi;
            transform_box(xf, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), lo, hi);
            BNode leaf;
            leaf.lo = lo; leaf.hi = hi; leaf.left = -1; leaf.right = -1;
            bn[i] = leaf;  // temporary position; moved to bn[ni + sorted position] after the sort
            const float cc[3] = { (lo.x + hi.x) * 0.5f, (lo.y + hi.y) * 0.5f, (lo.z + hi.z) * 0.5f };
            for (int k = 0; k < 3; k++) {
                bmin[k] = min(bmin[k], float_to_ordered(cc[k]));
                bmax[k] = max(bmax[k], float_to_ordered(cc[k]));
            }
        }
        for (int k = 0; k < 3; k++) {  // one shared-memory atomic per warp and component
            const uint32_t wmin = __reduce_min_sync(0xffffffffu, bmin[k]), wmax = __reduce_max_sync(0xffffffffu, bmax[k]);
            if ((tid & 31u) == 0u) {
                atomicMin(&s_bounds[k], wmin);
                atomicMax(&s_bounds[3 + k], wmax);
            }
        }
    }
    __syncthreads();
    TLAS_STAMP();  // 1: instance boxes
    // 2. Morton keys + bitonic sort of (key, index) in shared memory
    {
        const float3 lo = f3(ordered_to_float(s_bounds[0]), ordered_to_float(s_bounds[1]), ordered_to_float(s_bounds[2]));
        const float3 hi = f3(ordered_to_float(s_bounds[3]), ordered_to_float(s_bounds[4]), ordered_to_float(s_bounds[5]));
        const float3 ext = hi - lo;
        const float3 inv = f3(ext.x > 0.0f ? 1.0f / ext.x : 0.0f, ext.y > 0.0f ? 1.0f / ext.y : 0.0f, ext.z > 0.0f ? 1.0f / ext.z : 0.0f);
        for (uint32_t i = tid; i < p2; i += nt) {
            if (i < n) {
                // the top 53 bits of the 63-bit code (17.6 bits per axis) with the instance index below them: unique keys, so the
                // sort moves ONE 8-byte word per element and ties need no second array (instances whose codes agree that far
                // are ordered by index, as equal codes were before)
                const BNode &b = bn[i];
                static_assert(TLAS_FAST_MAX <= 2048, "11 index bits");
                keys[i] = ((morton63((b.lo + b.hi) * 0.5f, lo, inv) >> 10) << 11) | (uint64_t)i;
            } else keys[i] = ~0ull;  // padding sorts last
        }
    }
    __syncthreads();
    for (uint32_t k = 2; k <= p2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < p2; i += nt) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const uint64_t ka = keys[i], kb = keys[l];
                    const bool up = (i & k) == 0;
                    if ((ka > kb) == up) { keys[i] = kb; keys[l] = ka; }
                }
            }
            // partners closer than 32 sit in the same warp's 32 consecutive elements: between two such stages a warp barrier is
            // enough (35 of the 55 barriers at 1 024 keys)
            const uint32_t next_j = j > 1 ? (j >> 1) : k;
            if (j >= 32u || next_j >= 32u || (k == p2 && j == 1u)) __syncthreads(); else __syncwarp();
        }
    for (uint32_t i = tid; i < p2; i += nt) idx[i] = (uint32_t)(keys[i] & 2047u);
    __syncthreads();
    TLAS_STAMP();  // 2: keys + sort
    // 3. move the boxes to their sorted leaf slots.  Source bn[idx[i]] (i < n) and destination bn[ni + i] ranges overlap:
    //    stage through registers with a barrier in between (every thread owns at most ceil(n / nt) <= 2 leaves).
    {
        BNode tmp[(TLAS_FAST_MAX + (COOP ? TLAS_FAST_THREADS_COOP : TLAS_FAST_THREADS) - 1) / (COOP ? TLAS_FAST_THREADS_COOP : TLAS_FAST_THREADS)];
        int c = 0;
        for (uint32_t i = tid; i < n; i += nt) tmp[c++] = bn[idx[i]];
        __syncthreads();
        c = 0;
        for (uint32_t i = tid; i < n; i += nt) {
            bn[ni + i] = tmp[c++];
            count[ni + i] = 1;
            flags[i] = 0;
        }
    }
    __syncthreads();
    // 4. Karras links
    for (int i = (int)tid; i < ni; i += (int)nt) {
        int l, r, first, last;
        karras_node(keys, (int)n, i, l, r, first, last);
        bn[i].left = l;
        bn[i].right = r;
        parent[l] = (uint16_t)i;
        parent[r] = (uint16_t)i;
    }
    if (tid == 0) parent[0] = 0xffffu;
    __syncthreads();
    TLAS_STAMP();  // 3: leaf move + Karras links
    // 5. refit: the second thread to arrive at a node owns it
    for (uint32_t j = tid; j < n; j += nt) {
        uint32_t node = parent[ni + j];
        while (node != 0xffffu) {
            __threadfence_block();
            if (atomicAdd(&flags[node], 1u) == 0) break;
            __threadfence_block();
            const int l = bn[node].left, r = bn[node].right;
            const volatile BNode *vl = bn + l, *vr = bn + r;
            const float3 lo = fmin3(f3(vl->lo.x, vl->lo.y, vl->lo.z), f3(vr->lo.x, vr->lo.y, vr->lo.z));
            const float3 hi = fmax3(f3(vl->hi.x, vl->hi.y, vl->hi.z), f3(vr->hi.x, vr->hi.y, vr->hi.z));
            bn[node].lo = lo;
            bn[node].hi = hi;
            ((volatile int *)count)[node] = ((volatile int *)count)[l] + ((volatile int *)count)[r];
            node = parent[node];
        }
    }
    // 5b. optimal-collapse table (same dynamic program as k_bottom_up_dp; on the 1 000-instance lattice it is worth 14 % of
    //     trace throughput over the greedy collapse, the treelet pass nothing: 618 -> 704 Mrays/s)
    if (dp) {
        __syncthreads();
        TLAS_STAMP();  // 4: refit
        for (uint32_t i = tid; i < n; i += nt) flags[i] = 0;
        __syncthreads();
        for (uint32_t j = tid; j < n; j += nt) {
            {
                const BNode &leaf = bn[ni + j];
                DpEntry e;
                dp_leaf_entry(e, half_area(leaf.lo, leaf.hi), cost);
                dp[ni + j] = e;
            }
            uint32_t node = parent[ni + j];
            while (node != 0xffffu) {
                __threadfence_block();
                if (atomicAdd(&flags[node], 1u) == 0) break;
                __threadfence_block();
                const BNode b = bn[node];
                const DpEntry l = dp[b.left], r = dp[b.right];
                DpEntry e;
                dp_inner_entry(e, l, r, half_area(b.lo, b.hi), count[node], cost);
                dp[node] = e;
                node = parent[node];
            }
        }
    }
    if (tid == 0) {
        s_counters[0] = 1; s_counters[1] = 0; s_counters[2] = 0;
        s_items = 1; s_depth = 0;
        queue_a[0].bnode = 0; queue_a[0].wnode = 0;
    }
    __syncthreads();
    TLAS_STAMP();  // 5 (4 without the table): optimal-collapse table
    // 6. level-synchronous collapse
    for (;;) {
        const uint32_t n_items = s_items;
        if (n_items == 0) break;
        if (COOP) {
            for (uint32_t i = tid >> 5; i < n_items; i += nt >> 5)
                collapse_one_warp(bn, count, ni, queue_a[i], wide, &s_counters[0], &s_counters[1], idx, queue_b, &s_counters[2], leaf_prim, dp,
                                  s_expand[tid >> 5]);
        } else {
            for (uint32_t i = tid; i < n_items; i += nt)
                collapse_one(bn, count, ni, queue_a[i], wide, &s_counters[0], &s_counters[1], idx, nullptr, nullptr, queue_b, &s_counters[2],
                             leaf_prim, dp);
        }
        __syncthreads();
        if (tid == 0) { s_items = s_counters[2]; s_counters[2] = 0; s_depth++; }
        CollapseItem *t = queue_a; queue_a = queue_b; queue_b = t;
        __syncthreads();
        TLAS_STAMP();  // one per collapse level
    }
    // 7. instance leaf records in TLAS leaf order
    for (uint32_t j = tid; j < n; j += nt) {
        const uint32_t i = leaf_prim[j];
        float xf[16];
        const uint32_t blas = load_instance_placement(instances, i, xf);
        inst_leaves[j] = make_inst_leaf(xf, tlas_cap + blas, i);
    }
    if (tid == 0) {
        info->n_wide = s_counters[0]; info->depth = s_depth; info->leaf_count = s_counters[1];
        info->lo[0] = bn[0].lo.x; info->lo[1] = bn[0].lo.y; info->lo[2] = bn[0].lo.z;
        info->hi[0] = bn[0].hi.x; info->hi[1] = bn[0].hi.y; info->hi[2] = bn[0].hi.z;
        if (n_stamp < 16) info->t[n_stamp] = clock64();  // end
        for (int k = n_stamp + 1; k < 16; k++) info->t[k] = 0;
    }
#undef TLAS_STAMP
}

// ---------------------------------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

template <class T>
static cudaError_t dalloc(T **p, size_t count) { return cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)); }
// build scratch comes from the stream-ordered pool: no device-wide synchronisation per buffer, reused across rebuilds
// The pool is the ctx's PRIVATE one (solb_ctx_create): its raised release threshold keeps the scratch cached between rebuilds
// without touching the device's default pool, which belongs to the embedding application.
static thread_local cudaMemPool_t t_build_pool = nullptr;
template <class T>
static cudaError_t salloc(cudaStream_t st, T **p, size_t count) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    return t_build_pool ? cudaMallocFromPoolAsync((void **)p, bytes, t_build_pool, st) : cudaMallocAsync((void **)p, bytes, st);
}
void set_build_pool(cudaMemPool_t pool) { t_build_pool = pool; }

// SOLB_BUILD_TRACE=1: synchronise after every build phase and print its wall-clock share to stderr (debug aid)
struct PhaseTrace {
    cudaStream_t st;
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTrace(cudaStream_t s) : st(s) {
        const char *v = getenv("SOLB_BUILD_TRACE");
        on = v && *v == '1';
        if (on) { cudaStreamSynchronize(st); t0 = std::chrono::steady_clock::now(); }
    }
    void mark(const char *name) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[solb build] %-28s %9.3f ms\n", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Device scratch of one binary radix tree over n >= 2 primitives (Karras numbering: internal [0, n-2], leaf j -> n-1+j).
struct BinaryTree {
    uint32_t n = 0;
    uint64_t *keys = nullptr, *keys_tmp = nullptr;
    uint32_t *vals = nullptr, *vals_tmp = nullptr, *sort_scratch = nullptr, *flags = nullptr;
    BNode *bn = nullptr;
    int *parent = nullptr, *node_count = nullptr;
    float *node_cost = nullptr;
    DpEntry *dp = nullptr;  // optimal-collapse table (null: greedy collapse)
    cudaError_t alloc(cudaStream_t st, uint32_t n_) {
        n = n_;
        cudaError_t err = cudaSuccess;
        CK(salloc(st, &keys, n));
        CK(salloc(st, &keys_tmp, n));
        CK(salloc(st, &vals, n));
        CK(salloc(st, &vals_tmp, n));
        CK(salloc(st, &sort_scratch, onesweep_scratch_words(n)));
        CK(salloc(st, &bn, 2 * (size_t)n - 1));
        CK(salloc(st, &parent, 2 * (size_t)n - 1));
        CK(salloc(st, &node_count, 2 * (size_t)n - 1));
        CK(salloc(st, &node_cost, 2 * (size_t)n - 1));
        CK(salloc(st, &flags, n));
    done:
        return err;
    }
    void free(cudaStream_t st) {
        void *q[] = { keys, keys_tmp, vals, vals_tmp, sort_scratch, flags, bn, parent, node_count, node_cost, dp };
        for (void *x : q)
            if (x) cudaFreeAsync(x, st);
        *this = BinaryTree();
    }
};

// keys / vals filled: sort, link the radix tree, initialise the leaves
static cudaError_t tree_sort_and_link(cudaStream_t st, BinaryTree &T, const float4 *prim_lo, const float4 *prim_hi, int key_bits,
                                      uint64_t *launches) {
    cudaError_t err = radix_sort_pairs(st, T.keys, T.vals, T.keys_tmp, T.vals_tmp, T.n, key_bits, T.sort_scratch, launches);
    if (err != cudaSuccess) return err;
    k_hierarchy<<<(T.n + 255) / 256, 256, 0, st>>>(T.keys, T.vals, prim_lo, prim_hi, (int)T.n, T.bn, T.parent, T.node_count, T.node_cost,
                                                     T.flags);
    *launches += 1;
    // the sort's ping-pong buffers are dead from here on: hand them back so later scratch (collapse table, queues) reuses them
    cudaFreeAsync(T.keys_tmp, st); cudaFreeAsync(T.vals_tmp, st); cudaFreeAsync(T.sort_scratch, st);
    T.keys_tmp = nullptr; T.vals_tmp = nullptr; T.sort_scratch = nullptr;
    return cudaGetLastError();
}

// keys / vals filled: sort, then PLOC over the sorted leaves.  Produces everything tree_sort_and_link + tree_refit_optimize do
// (binary nodes with boxes, parents, counts, SAH cost, optimal-collapse table), root = node 0.
static cudaError_t tree_sort_and_ploc(cudaStream_t st, BinaryTree &T, const float4 *prim_lo, const float4 *prim_hi, int key_bits,
                                      const BuildOptions &opt, uint32_t *h_count /* pinned */, uint64_t *launches, uint32_t *iterations) {
    cudaError_t err = radix_sort_pairs(st, T.keys, T.vals, T.keys_tmp, T.vals_tmp, T.n, key_bits, T.sort_scratch, launches);
    if (err != cudaSuccess) return err;
    const uint32_t n = T.n;
    float4 *c_lo = nullptr, *c_hi = nullptr, *t_lo = nullptr, *t_hi = nullptr;
    uint32_t *block_counts = nullptr, *offsets = nullptr, *dev_count = nullptr;
    const uint32_t max_blocks = (n + PLOC_CH - 1) / PLOC_CH;
    // the sort's ping-pong buffers and the keys are dead from here on
    cudaFreeAsync(T.keys_tmp, st); cudaFreeAsync(T.vals_tmp, st); cudaFreeAsync(T.sort_scratch, st); cudaFreeAsync(T.keys, st);
    T.keys_tmp = nullptr; T.vals_tmp = nullptr; T.sort_scratch = nullptr; T.keys = nullptr;
    DpCost cost;
    cost.cn = opt.dp_cn; cost.cp = opt.dp_cp; cost.max_leaf = opt.dp_max_leaf;
    CK(salloc(st, &c_lo, n));
    CK(salloc(st, &c_hi, n));
    CK(salloc(st, &t_lo, (size_t)max_blocks * PLOC_CH));
    CK(salloc(st, &t_hi, (size_t)max_blocks * PLOC_CH));
    CK(salloc(st, &block_counts, max_blocks));
    CK(salloc(st, &offsets, max_blocks));
    CK(salloc(st, &dev_count, 2));
    if (opt.dp_collapse) CK(salloc(st, &T.dp, 2 * (size_t)n - 1));
    CK(cudaMemsetAsync(dev_count, 0, 2 * sizeof(uint32_t), st));  // [0] cluster count after the iteration, [1] nodes created
    k_ploc_init<<<(n + 255) / 256, 256, 0, st>>>((int)n, T.vals, prim_lo, prim_hi, T.bn, T.parent, T.node_count, T.node_cost, T.dp, cost,
                                                  c_lo, c_hi);
    *launches += 1;
    {
        uint32_t count = n, iters = 0;
        while (count > 1) {
            const uint32_t nb = (count + PLOC_CH - 1) / PLOC_CH;
            k_ploc_merge<<<nb, PLOC_CH, 0, st>>>((int)n, (int)count, c_lo, c_hi, t_lo, t_hi, block_counts, dev_count + 1, T.bn, T.parent,
                                                 T.node_count, T.node_cost, T.dp, cost);
            k_ploc_scan<<<1, 1024, 0, st>>>(block_counts, nb, offsets, dev_count);
            k_ploc_scatter<<<nb, PLOC_CH, 0, st>>>(t_lo, t_hi, block_counts, offsets, c_lo, c_hi);
            *launches += 3;
            CK(cudaMemcpyAsync(h_count, dev_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (*h_count >= count || *h_count == 0) { err = cudaErrorUnknown; goto done; }  // every iteration merges at least one pair
            count = *h_count;
            if (++iters > 4096) { err = cudaErrorUnknown; goto done; }
        }
        if (iterations) *iterations = iters;
    }
    CK(cudaGetLastError());
done:
    {
        void *q[] = { c_lo, c_hi, t_lo, t_hi, block_counts, offsets, dev_count };
        for (void *x : q)
            if (x) cudaFreeAsync(x, st);
    }
    return err;
}

// refit (boxes, counts, SAH cost) and treelet restructuring; walks stop at nodes whose parent is -1
static cudaError_t tree_refit_optimize(cudaStream_t st, BinaryTree &T, const BuildOptions &opt, int treelet_passes, float *sah_lbvh_out,
                                       uint64_t *launches) {
    const uint32_t n = T.n, nb = (n + 255) / 256;
    cudaError_t err = cudaSuccess;
    k_bottom_up<<<nb, 256, 0, st>>>((int)n, T.bn, T.parent, T.node_count, T.node_cost, T.flags, 0, 0);
    *launches += 1;
    if (sah_lbvh_out) CK(cudaMemcpyAsync(sah_lbvh_out, T.node_cost, sizeof(float), cudaMemcpyDeviceToHost, st));
    for (int pass = 0; pass < treelet_passes; pass++) {
        k_clear_u32<<<nb, 256, 0, st>>>(T.flags, n);
        if (opt.coop_treelet)
            k_bottom_up_coop<<<(n + TL_BLOCK - 1) / TL_BLOCK, TL_BLOCK, 0, st>>>((int)n, T.bn, T.parent, T.node_count, T.node_cost, T.flags,
                                                                                  opt.treelet_gamma);
        else
            k_bottom_up<<<(n + 63) / 64, 64, 0, st>>>((int)n, T.bn, T.parent, T.node_count, T.node_cost, T.flags, 1, opt.treelet_gamma);
        *launches += 2;
    }
    if (opt.dp_collapse) {
        CK(salloc(st, &T.dp, 2 * (size_t)n - 1));
        k_clear_u32<<<nb, 256, 0, st>>>(T.flags, n);
        DpCost cost;
        cost.cn = opt.dp_cn; cost.cp = opt.dp_cp; cost.max_leaf = opt.dp_max_leaf;
        k_bottom_up_dp<<<nb, 256, 0, st>>>((int)n, T.bn, T.parent, T.node_count, T.flags, T.dp, cost);
        *launches += 2;
    }
    CK(cudaGetLastError());
done:
    return err;
}

// Level-synchronous collapse.  queue_a holds n_items seeds; counters = {wide nodes allocated, leaf slots allocated, 0}
// must already be on the device.  On return h_counters holds the final counts and *levels the number of levels.
static cudaError_t collapse_levels(cudaStream_t st, const BinaryTree &T, CollapseItem *queue_a, CollapseItem *queue_b, uint32_t n_items,
                                   Node8 *wide, uint32_t *counters, const Tri48 *tri_src, Tri48 *tri_out, uint32_t *leaf_prim_out,
                                   uint32_t h_counters[3], uint32_t *levels, uint64_t *launches) {
    cudaError_t err = cudaSuccess;
    uint32_t depth = 0;
    while (n_items) {
        k_collapse_level<<<(n_items + 63) / 64, 64, 0, st>>>(T.bn, T.node_count, (int)T.n - 1, queue_a, n_items, wide, counters, T.vals,
                                                             tri_src, tri_out, queue_b, leaf_prim_out, T.dp);
        *launches += 1;
        CK(cudaMemcpyAsync(h_counters, counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        n_items = h_counters[2];
        h_counters[2] = 0;
        CK(cudaMemcpyAsync(counters + 2, &h_counters[2], sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        std::swap(queue_a, queue_b);
        depth++;
        if (depth > SOLB_MAX_WIDE_DEPTH) { err = cudaErrorLaunchOutOfResources; goto done; }
    }
    *levels = depth;
done:
    return err;
}

cudaError_t build_accel(cudaStream_t st, const DeviceSceneView &sv, AccelStorage &out, const BuildOptions &opt, uint64_t *launches) {
    cudaError_t err = cudaSuccess;
    const uint32_t n = sv.n_tris;
    Tri48 *tri_world = nullptr;
    float4 *prim_lo = nullptr, *prim_hi = nullptr;
    uint32_t *bounds = nullptr, *counters = nullptr;
    BinaryTree T;
    CollapseItem *queue_a = nullptr, *queue_b = nullptr;
    Node8 *wide = nullptr;
    Tri48 *tri_out = nullptr;
    const uint32_t nb = (n + 255) / 256;
    uint32_t h_counters[3];
    uint32_t depth = 0;

    PhaseTrace trace(st);
    // A rebuild (moved instances of a flattened scene: every frame of a dynamic scene) reuses the node and triangle arrays of
    // the structure it replaces when they are large enough: cudaFree + cudaMalloc of the ~1.2 GB of a 20 M-triangle scene are
    // ~5 ms of driver time, a fifth of the whole rebuild (profiles/r02_ploc_build.txt, "alloc + prep").
    Node8 *old_nodes = out.two_level ? nullptr : out.nodes;
    Tri48 *old_tris = out.two_level ? nullptr : out.tris;
    const size_t old_nodes_cap = out.nodes_cap, old_tris_cap = out.tris_cap;
    size_t tri_cap = 0;
    if (old_nodes) out.nodes = nullptr;
    if (old_tris) out.tris = nullptr;
    out.release();
    out.n_tris = n;
    CK(salloc(st, &wide, std::max<uint32_t>(n, 1)));
    if (old_tris && old_tris_cap >= std::max<size_t>(n, 1)) {
        tri_out = old_tris;
        tri_cap = old_tris_cap;
    } else {
        cudaFree(old_tris);
        tri_cap = std::max<size_t>(n, 1);
        CK(dalloc(&tri_out, n));
    }
    old_tris = nullptr;
    if (n == 0) {
        k_empty_root<<<1, 1, 0, st>>>(wide);
        *launches += 1;
        out.n_wide = 1;
        out.depth = 1;
        goto finish;
    }
    CK(salloc(st, &tri_world, n));
    CK(salloc(st, &prim_lo, n));
    CK(salloc(st, &prim_hi, n));
    CK(salloc(st, &bounds, 8));
    k_init_bounds<<<1, 32, 0, st>>>(bounds);
    k_prep_tris<<<nb, 256, 0, st>>>(sv, tri_world, prim_lo, prim_hi, bounds);
    *launches += 2;
    if (n == 1) {
        k_single_tri_root<<<1, 1, 0, st>>>(prim_lo, prim_hi, tri_world, wide, tri_out);
        *launches += 1;
        out.n_wide = 1;
        out.depth = 1;
        goto finish;
    }
    trace.mark("alloc + prep");
    CK(T.alloc(st, n));
    k_morton<<<nb, 256, 0, st>>>(prim_lo, prim_hi, n, bounds, T.keys, T.vals);
    *launches += 1;
    trace.mark("tree alloc + morton");
    if (opt.ploc) {
        uint32_t *h_count = nullptr, iters = 0;
        CK(cudaMallocHost((void **)&h_count, sizeof(uint32_t)));
        err = tree_sort_and_ploc(st, T, prim_lo, prim_hi, 63, opt, h_count, launches, &iters);
        cudaFreeHost(h_count);
        CK(err);
        cudaFreeAsync(prim_lo, st); cudaFreeAsync(prim_hi, st);  // leaf boxes live in the tree now
        prim_lo = prim_hi = nullptr;
        trace.mark("sort + PLOC");
        CK(cudaMemcpyAsync(&out.sah_lbvh, T.node_cost, sizeof(float), cudaMemcpyDeviceToHost, st));  // (no LBVH stage to compare with)
    } else {
        CK(tree_sort_and_link(st, T, prim_lo, prim_hi, 63, launches));
        cudaFreeAsync(prim_lo, st); cudaFreeAsync(prim_hi, st); cudaFreeAsync(T.keys, st);  // leaf boxes live in the tree now
        prim_lo = prim_hi = nullptr; T.keys = nullptr;
        trace.mark("sort + hierarchy");
        CK(tree_refit_optimize(st, T, opt, opt.treelet_passes, &out.sah_lbvh, launches));
    }
    CK(cudaMemcpyAsync(&out.sah_final, T.node_cost, sizeof(float), cudaMemcpyDeviceToHost, st));
    {
        BNode root;
        CK(cudaMemcpyAsync(&root, T.bn, sizeof(BNode), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const float a = half_area(root.lo, root.hi);
        if (a > 0.0f) { out.sah_lbvh /= a; out.sah_final /= a; }
        out.lo[0] = root.lo.x; out.lo[1] = root.lo.y; out.lo[2] = root.lo.z;
        out.hi[0] = root.hi.x; out.hi[1] = root.hi.y; out.hi[2] = root.hi.z;
    }
    trace.mark("refit + treelets");
    // collapse, one launch per level of the wide tree
    cudaFreeAsync(T.node_cost, st); cudaFreeAsync(T.parent, st); cudaFreeAsync(T.flags, st);
    T.node_cost = nullptr; T.parent = nullptr; T.flags = nullptr;
    CK(salloc(st, &queue_a, n));
    CK(salloc(st, &queue_b, n));
    CK(salloc(st, &counters, 4));
    {
        CollapseItem rootItem;
        rootItem.bnode = 0;
        rootItem.wnode = 0;
        CK(cudaMemcpyAsync(queue_a, &rootItem, sizeof(rootItem), cudaMemcpyHostToDevice, st));
        h_counters[0] = 1; h_counters[1] = 0; h_counters[2] = 0;
        CK(cudaMemcpyAsync(counters, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st));
        CK(collapse_levels(st, T, queue_a, queue_b, 1, wide, counters, tri_world, tri_out, nullptr, h_counters, &depth, launches));
        out.n_wide = h_counters[0];
        out.depth = depth;
        if (h_counters[1] != n) { err = cudaErrorUnknown; goto done; }  // every triangle must land in exactly one leaf
    }
    trace.mark("collapse");
finish:
    CK(cudaStreamSynchronize(st));
    // shrink the node array to its final size
    {
        Node8 *final_nodes = nullptr;
        if (old_nodes && old_nodes_cap >= out.n_wide) {
            final_nodes = old_nodes;
            out.nodes_cap = old_nodes_cap;
        } else {
            cudaFree(old_nodes);
            old_nodes = nullptr;
            out.nodes_cap = (size_t)out.n_wide + out.n_wide / 16;  // some slack: a rebuild after small moves rarely has the same count
            CK(dalloc(&final_nodes, out.nodes_cap));
        }
        old_nodes = nullptr;
        CK(cudaMemcpyAsync(final_nodes, wide, sizeof(Node8) * out.n_wide, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        out.nodes = final_nodes;
        out.tris = tri_out;
        out.tris_cap = tri_cap;
        tri_out = nullptr;
        out.n_binary = n >= 2 ? 2 * n - 1 : n;
    }
    trace.mark("final node copy");
done:
    {
        void *scratch[] = { tri_world, prim_lo, prim_hi, bounds, counters, queue_a, queue_b, wide };
        for (void *q : scratch)
            if (q) cudaFreeAsync(q, st);
        T.free(st);
    }
    cudaFree(tri_out);
    cudaFree(old_nodes);  // (error paths only: consumed or freed above otherwise)
    cudaFree(old_tris);
    trace.mark("free scratch");
    if (err != cudaSuccess) out.release();
    return err;
}

// ---------------------------------------------------------------------------------------------------
static bool tlas_fast_enabled() {
    const char *v = getenv("SOLB_TLAS_FAST");
    return !(v && *v == '0');
}

// TLAS over the instances' world boxes, written into nodes[0, tlas_cap) and inst_leaves (both preallocated by
// build_accel_two_level).  This is all TLAS::regenerate has to redo when instance transforms change.
cudaError_t rebuild_tlas(cudaStream_t st, const DeviceSceneView &sv, AccelStorage &out, uint64_t *launches) {
    cudaError_t err = cudaSuccess;
    const uint32_t n = sv.n_instances;
    float4 *prim_lo = nullptr, *prim_hi = nullptr;
    uint32_t *bounds = nullptr, *counters = nullptr, *leaf_prim = nullptr;
    BinaryTree T;
    CollapseItem *queue_a = nullptr, *queue_b = nullptr;
    uint32_t h_counters[3];
    uint32_t depth = 1;
    BuildOptions opt;
    int tlas_passes = 2;
    if (const char *v = getenv("SOLB_TLAS_TREELET_PASSES")) tlas_passes = std::max(0, std::min(8, atoi(v)));
    if (const char *v = getenv("SOLB_TLAS_DP")) opt.dp_collapse = atoi(v) != 0;
    // TLAS cost model: an instance entry is a whole BLAS walk, not a triangle test
    opt.dp_cp = 8.0f;
    opt.dp_max_leaf = 1;
    if (const char *v = getenv("SOLB_TLAS_CP")) opt.dp_cp = (float)atof(v);
    if (const char *v = getenv("SOLB_TLAS_MAX_LEAF")) opt.dp_max_leaf = std::max(1, std::min(SOLB_MAX_LEAF_TRIS, atoi(v)));
    DpCost tlas_cost;
    tlas_cost.cn = opt.dp_cn; tlas_cost.cp = opt.dp_cp; tlas_cost.max_leaf = opt.dp_max_leaf;
    if (!out.two_level || !out.nodes || n > out.tlas_cap) return cudaErrorInvalidValue;
    out.sah_lbvh = out.sah_final = 0.0f;
    if (n == 0) {
        k_empty_root<<<1, 1, 0, st>>>(out.nodes);
        *launches += 1;
        out.n_tlas_wide = 1;
        goto finish;
    }
    if (n >= 2 && n <= TLAS_FAST_MAX && tlas_fast_enabled()) {
        // per device and cheap: set on every call rather than caching a process-wide flag (one process may own several devices)
        if (!out.d_tlas_info) {
            CK(cudaMalloc(&out.d_tlas_info, sizeof(TlasFastInfo)));
            CK(cudaMallocHost(&out.h_tlas_info, sizeof(TlasFastInfo)));
            CK(cudaMalloc(&out.d_tlas_dp, sizeof(DpEntry) * 2 * (size_t)out.tlas_cap));
        }
        const char *coop_env = getenv("SOLB_TLAS_COOP");
        const bool coop = opt.dp_collapse && !(coop_env && *coop_env == '0');
        if (coop) {
            const bool dp_smem = tlas_fast_smem_bytes_dp(n) <= TLAS_FAST_SMEM_LIMIT;
            const size_t bytes = dp_smem ? tlas_fast_smem_bytes_dp(n) : tlas_fast_smem_bytes(n);
            CK(cudaFuncSetAttribute(k_tlas_build_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TLAS_FAST_SMEM_LIMIT));
            k_tlas_build_small<true><<<1, TLAS_FAST_THREADS_COOP, bytes, st>>>(sv.instances, n, out.blas_box, out.tlas_cap, out.nodes,
                                                                              out.inst_leaves, (TlasFastInfo *)out.d_tlas_info,
                                                                              dp_smem ? nullptr : (DpEntry *)out.d_tlas_dp, dp_smem ? 1 : 0,
                                                                              tlas_cost);
        } else {
            CK(cudaFuncSetAttribute(k_tlas_build_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tlas_fast_smem_bytes(TLAS_FAST_MAX)));
            k_tlas_build_small<false><<<1, TLAS_FAST_THREADS, tlas_fast_smem_bytes(n), st>>>(sv.instances, n, out.blas_box, out.tlas_cap, out.nodes,
                                                                                            out.inst_leaves, (TlasFastInfo *)out.d_tlas_info,
                                                                                            opt.dp_collapse ? (DpEntry *)out.d_tlas_dp : nullptr, 0,
                                                                                            tlas_cost);
        }
        *launches += 1;
        CK(cudaMemcpyAsync(out.h_tlas_info, out.d_tlas_info, sizeof(TlasFastInfo), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const TlasFastInfo h_info = *(const TlasFastInfo *)out.h_tlas_info;
        if (const char *v = getenv("SOLB_TLAS_TRACE")) {
            if (atoi(v)) {
                fprintf(stderr, "[solb tlas] n = %u, cycles per phase:", n);
                for (int k = 1; k < 16 && h_info.t[k]; k++) fprintf(stderr, " %lld", h_info.t[k] - h_info.t[k - 1]);
                fprintf(stderr, "\n");
            }
        }
        if (h_info.leaf_count != n || h_info.n_wide > out.tlas_cap) { err = cudaErrorUnknown; goto done; }
        out.n_tlas_wide = h_info.n_wide;
        depth = h_info.depth;
        for (int k = 0; k < 3; k++) { out.lo[k] = h_info.lo[k]; out.hi[k] = h_info.hi[k]; }
        goto finish;
    }
    CK(salloc(st, &prim_lo, n));
    CK(salloc(st, &prim_hi, n));
    CK(salloc(st, &bounds, 8));
    k_init_bounds<<<1, 32, 0, st>>>(bounds);
    k_inst_boxes<<<(n + 127) / 128, 128, 0, st>>>(sv.instances, n, out.blas_box, prim_lo, prim_hi, bounds);
    *launches += 2;
    if (n == 1) {
        k_single_inst_root<<<1, 1, 0, st>>>(prim_lo, prim_hi, out.nodes);
        k_write_inst_leaves<<<1, 32, 0, st>>>(nullptr, 1, sv.instances, out.tlas_cap, out.inst_leaves);
        *launches += 2;
        out.n_tlas_wide = 1;
        float4 h[2];
        CK(cudaMemcpyAsync(&h[0], prim_lo, sizeof(float4), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&h[1], prim_hi, sizeof(float4), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        out.lo[0] = h[0].x; out.lo[1] = h[0].y; out.lo[2] = h[0].z;
        out.hi[0] = h[1].x; out.hi[1] = h[1].y; out.hi[2] = h[1].z;
        goto finish;
    }
    CK(T.alloc(st, n));
    k_morton<<<(n + 255) / 256, 256, 0, st>>>(prim_lo, prim_hi, n, bounds, T.keys, T.vals);
    *launches += 1;
    CK(tree_sort_and_link(st, T, prim_lo, prim_hi, 63, launches));
    CK(tree_refit_optimize(st, T, opt, tlas_passes, &out.sah_lbvh, launches));
    CK(cudaMemcpyAsync(&out.sah_final, T.node_cost, sizeof(float), cudaMemcpyDeviceToHost, st));
    {
        BNode root;
        CK(cudaMemcpyAsync(&root, T.bn, sizeof(BNode), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const float a = half_area(root.lo, root.hi);
        if (a > 0.0f) { out.sah_lbvh /= a; out.sah_final /= a; }
        out.lo[0] = root.lo.x; out.lo[1] = root.lo.y; out.lo[2] = root.lo.z;
        out.hi[0] = root.hi.x; out.hi[1] = root.hi.y; out.hi[2] = root.hi.z;
    }
    CK(salloc(st, &queue_a, n));
    CK(salloc(st, &queue_b, n));
    CK(salloc(st, &counters, 4));
    CK(salloc(st, &leaf_prim, n));
    {
        CollapseItem rootItem;
        rootItem.bnode = 0;
        rootItem.wnode = 0;
        CK(cudaMemcpyAsync(queue_a, &rootItem, sizeof(rootItem), cudaMemcpyHostToDevice, st));
        h_counters[0] = 1; h_counters[1] = 0; h_counters[2] = 0;
        CK(cudaMemcpyAsync(counters, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st));
        CK(collapse_levels(st, T, queue_a, queue_b, 1, out.nodes, counters, nullptr, nullptr, leaf_prim, h_counters, &depth, launches));
        if (h_counters[1] != n || h_counters[0] > out.tlas_cap) { err = cudaErrorUnknown; goto done; }
        out.n_tlas_wide = h_counters[0];
        k_write_inst_leaves<<<(n + 127) / 128, 128, 0, st>>>(leaf_prim, n, sv.instances, out.tlas_cap, out.inst_leaves);
        *launches += 1;
    }
finish:
    out.tlas_depth = depth;
    out.depth = out.tlas_depth + out.blas_depth;
    if (out.depth > SOLB_MAX_WIDE_DEPTH) { err = cudaErrorLaunchOutOfResources; goto done; }
    CK(cudaGetLastError());
done:
    {
        void *scratch[] = { prim_lo, prim_hi, bounds, counters, leaf_prim, queue_a, queue_b };
        for (void *q : scratch)
            if (q) cudaFreeAsync(q, st);
        T.free(st);
    }
    return err;
}

cudaError_t build_accel_two_level(cudaStream_t st, const DeviceSceneView &sv, AccelStorage &out, const BuildOptions &opt,
                                  uint64_t *launches) {
    cudaError_t err = cudaSuccess;
    const uint32_t n = sv.n_geom_tris, n_blas = sv.n_blas, n_inst = sv.n_instances;
    const uint32_t tlas_cap = std::max<uint32_t>(n_inst, 1);
    Tri48 *tri_obj = nullptr, *tri_out = nullptr;
    float4 *prim_lo = nullptr, *prim_hi = nullptr;
    uint32_t *blas_bounds = nullptr, *counters = nullptr, *d_list = nullptr;
    int *blas_root = nullptr;
    BinaryTree T;
    CollapseItem *queue_a = nullptr, *queue_b = nullptr;
    Node8 *wide = nullptr;
    const uint32_t nb = (n + 255) / 256;
    uint32_t h_counters[3] = { tlas_cap + n_blas, 0, 0 };
    uint32_t depth = 1;
    std::vector<uint32_t> h_list;  // multi-triangle BLASes first, then single-triangle ones
    uint32_t n_multi = 0, n_single = 0;
    std::vector<DeviceBlas> h_blas(n_blas);

    PhaseTrace trace(st);
    out.release();
    out.two_level = true;
    out.n_tris = n;
    out.n_blas = n_blas;
    out.tlas_cap = tlas_cap;
    // worst case wide-node count: TLAS region + one root per BLAS + (#triangles - 1) interiors
    CK(salloc(st, &wide, (size_t)tlas_cap + n_blas + n));
    CK(dalloc(&tri_out, n));
    CK(dalloc(&out.inst_leaves, n_inst));
    CK(dalloc(&out.blas_box, 2 * (size_t)n_blas));
    if (n_blas) {
        CK(cudaMemcpyAsync(h_blas.data(), sv.blas, n_blas * sizeof(DeviceBlas), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (uint32_t b = 0; b < n_blas; b++)
            if (h_blas[b].n_indices / 3 >= 2) h_list.push_back(b);
        n_multi = (uint32_t)h_list.size();
        for (uint32_t b = 0; b < n_blas; b++)
            if (h_blas[b].n_indices / 3 == 1) h_list.push_back(b);
        n_single = (uint32_t)h_list.size() - n_multi;
        if (n_multi + n_single != n_blas) { err = cudaErrorInvalidValue; goto done; }  // empty BLASes are rejected at scene creation
        CK(salloc(st, &d_list, n_blas));
        CK(cudaMemcpyAsync(d_list, h_list.data(), n_blas * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CK(salloc(st, &tri_obj, n));
        CK(salloc(st, &prim_lo, n));
        CK(salloc(st, &prim_hi, n));
        CK(salloc(st, &blas_bounds, 6 * (size_t)n_blas));
        CK(salloc(st, &blas_root, n_blas));
        k_init_bounds_n<<<(6 * n_blas + 255) / 256, 256, 0, st>>>(blas_bounds, n_blas);
        k_prep_tris_obj<<<nb, 256, 0, st>>>(sv, tri_obj, prim_lo, prim_hi, blas_bounds);
        *launches += 2;
        trace.mark("2L alloc + prep");
        if (n >= 2) {
            int id_bits = 0;
            while ((1ull << id_bits) < n_blas) id_bits++;
            const int morton_bits = ((63 - id_bits) / 3) * 3;
            CK(T.alloc(st, n));
            k_morton_seg<<<nb, 256, 0, st>>>(prim_lo, prim_hi, n, blas_bounds, morton_bits, T.keys, T.vals);
            *launches += 1;
            trace.mark("2L tree alloc + morton");
            CK(tree_sort_and_link(st, T, prim_lo, prim_hi, morton_bits + id_bits, launches));
            k_blas_roots<<<(n_blas + 127) / 128, 128, 0, st>>>(T.keys, (int)n, sv.blas, n_blas, T.parent, blas_root);
            *launches += 1;
            trace.mark("2L sort + hierarchy + roots");
            CK(tree_refit_optimize(st, T, opt, opt.treelet_passes, nullptr, launches));
            k_blas_boxes<<<(n_blas + 127) / 128, 128, 0, st>>>(T.bn, blas_root, n_blas, out.blas_box);
            *launches += 1;
            trace.mark("2L refit + treelets");
        } else {  // one BLAS holding one triangle
            CK(cudaMemcpyAsync(out.blas_box, prim_lo, sizeof(float4), cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(out.blas_box + 1, prim_hi, sizeof(float4), cudaMemcpyDeviceToDevice, st));
        }
        CK(salloc(st, &counters, 4));
        h_counters[1] = n_single;  // triangle slots [0, n_single) belong to the single-triangle BLASes
        CK(cudaMemcpyAsync(counters, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st));
        if (n_single) {
            k_single_tri_blas<<<(n_single + 127) / 128, 128, 0, st>>>(d_list + n_multi, n_single, sv.blas, prim_lo, prim_hi, tri_obj, tlas_cap,
                                                                      wide, tri_out);
            *launches += 1;
        }
        if (n_multi) {
            CK(salloc(st, &queue_a, n));
            CK(salloc(st, &queue_b, n));
            k_seed_blas_collapse<<<(n_multi + 127) / 128, 128, 0, st>>>(d_list, n_multi, blas_root, tlas_cap, queue_a);
            *launches += 1;
            CK(collapse_levels(st, T, queue_a, queue_b, n_multi, wide, counters, tri_obj, tri_out, nullptr, h_counters, &depth, launches));
            if (h_counters[1] != n) { err = cudaErrorUnknown; goto done; }
        }
    }
    trace.mark("2L collapse");
    out.blas_depth = depth;
    out.n_wide = h_counters[0];
    CK(cudaStreamSynchronize(st));
    {
        Node8 *final_nodes = nullptr;
        CK(dalloc(&final_nodes, out.n_wide));
        // the TLAS region is written by rebuild_tlas below (unused slots stay zero = empty nodes); copy the BLAS part
        CK(cudaMemsetAsync(final_nodes, 0, sizeof(Node8) * std::min(out.n_wide, tlas_cap), st));
        if (out.n_wide > tlas_cap)
            CK(cudaMemcpyAsync(final_nodes + tlas_cap, wide + tlas_cap, sizeof(Node8) * (out.n_wide - tlas_cap), cudaMemcpyDeviceToDevice, st));
        out.nodes = final_nodes;
        out.tris = tri_out;
        tri_out = nullptr;
        out.n_binary = n >= 2 ? 2 * n - 1 : n;
    }
    trace.mark("2L final node copy");
    CK(rebuild_tlas(st, sv, out, launches));
    CK(cudaStreamSynchronize(st));
    trace.mark("2L tlas");
done:
    {
        void *scratch[] = { tri_obj, prim_lo, prim_hi, blas_bounds, counters, d_list, blas_root, queue_a, queue_b, wide };
        for (void *q : scratch)
            if (q) cudaFreeAsync(q, st);
        T.free(st);
    }
    cudaFree(tri_out);
    if (err != cudaSuccess) out.release();
    return err;
}

// test hook: sort (keys, vals) of n elements already on the device
cudaError_t sort_pairs_device(cudaStream_t st, uint64_t *keys, uint32_t *vals, uint32_t n, int key_bits) {
    uint64_t *kt = nullptr;
    uint32_t *vt = nullptr, *th = nullptr;
    uint64_t launches = 0;
    cudaError_t err = cudaSuccess;
    if (n == 0) return cudaSuccess;
    CK(dalloc(&kt, n));
    CK(dalloc(&vt, n));
    CK(dalloc(&th, onesweep_scratch_words(n)));
    CK(radix_sort_pairs(st, keys, vals, kt, vt, n, key_bits, th, &launches));
    CK(cudaStreamSynchronize(st));
done:
    cudaFree(kt); cudaFree(vt); cudaFree(th);
    return err;
}

}  // namespace solb
