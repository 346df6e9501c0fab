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
        t.v2 = make_float4(p[2].x, p[2].y, p[2].z, __uint_as_float(g));
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
// leaves with atomic flags, but when lanes of a warp reach treelet roots the WHOLE warp optimises those
// treelets one after another: 32 lanes share the 2^7 subset areas and the dynamic program over subset
// sizes, so one treelet costs ~1 us instead of the ~50 us of the per-thread version (k_bottom_up mode 1),
// which also shortens the serial critical path up the tree by the same factor.
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

__device__ void coop_optimize_treelet(BNode *bn, int *parent, float *node_cost, int *node_count, int n_internal, int root,
                                      WarpTreeletScratch &ws, int lane) {
    if (lane == 0) {  // treelet formation: expand the largest-area internal leaf until 7 leaves
        int nl = 0, ni = 0;
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
        ws.nl = nl;
        ws.changed = 0;
    }
    __syncwarp();
    const int nl = ws.nl;
    if (nl < 3) return;
    const int full = (1 << nl) - 1;
    if (lane < nl) {
        const BNode b = bn[ws.leaves[lane]];
        ws.lo[lane][0] = b.lo.x; ws.lo[lane][1] = b.lo.y; ws.lo[lane][2] = b.lo.z;
        ws.hi[lane][0] = b.hi.x; ws.hi[lane][1] = b.hi.y; ws.hi[lane][2] = b.hi.z;
        ws.lcost[lane] = node_cost[ws.leaves[lane]];
        ws.lcount[lane] = node_count[ws.leaves[lane]];
    }
    __syncwarp();
    for (int s = lane + 1; s <= full; s += 32) {  // subset areas and "fits in a leaf" flags
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
    for (int k = 2; k <= nl; k++) {  // dynamic program over subset sizes
        for (int s = lane + 1; s <= full; s += 32) {
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
    if (lane == 0 && ws.cost[full] < node_cost[root] * 0.9999f) {  // rebuild the topology, reusing the internal node ids
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
    __shared__ WarpTreeletScratch scratch[TL_BLOCK / 32];
    WarpTreeletScratch &ws = scratch[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
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
        while (m) {
            const int leader = __ffs(m) - 1;
            const int root = __shfl_sync(0xffffffffu, node, leader);
            coop_optimize_treelet(bn, parent, node_cost, node_count, n - 1, root, ws, lane);
            m &= m - 1;
        }
        if (active) {
            node = parent[node];
            if (node < 0) active = false;
        }
    }
}

__global__ void k_clear_u32(uint32_t *p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

__global__ void k_collapse_level(const BNode *bn, const int *node_count, int n_internal, const CollapseItem *queue_in,
                                 uint32_t n_items, Node8 *wide, uint32_t *counters /*0: wide, 1: tris, 2: queue_out*/,
                                 const uint32_t *sorted_prim, const Tri48 *tri_world, Tri48 *tri_out, CollapseItem *queue_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    collapse_one(bn, node_count, n_internal, queue_in[i], wide, &counters[0], &counters[1], sorted_prim, tri_world, tri_out,
                 queue_out, &counters[2]);
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
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

template <class T>
static cudaError_t dalloc(T **p, size_t count) { return cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)); }
// build scratch comes from the stream-ordered pool: no device-wide synchronisation per buffer, reused across rebuilds
template <class T>
static cudaError_t salloc(cudaStream_t st, T **p, size_t count) {
    return cudaMallocAsync((void **)p, std::max<size_t>(count, 1) * sizeof(T), st);
}

cudaError_t build_accel(cudaStream_t st, const DeviceSceneView &sv, AccelStorage &out, const BuildOptions &opt, uint64_t *launches) {
    cudaError_t err = cudaSuccess;
    const uint32_t n = sv.n_tris;
    Tri48 *tri_world = nullptr;
    float4 *prim_lo = nullptr, *prim_hi = nullptr;
    uint32_t *bounds = nullptr, *vals = nullptr, *vals_tmp = nullptr, *tile_hist = nullptr, *flags = nullptr, *counters = nullptr;
    uint64_t *keys = nullptr, *keys_tmp = nullptr;
    BNode *bn = nullptr;
    int *parent = nullptr, *node_count = nullptr;
    float *node_cost = nullptr;
    CollapseItem *queue_a = nullptr, *queue_b = nullptr;
    Node8 *wide = nullptr;
    Tri48 *tri_out = nullptr;
    const uint32_t T = 256;
    const uint32_t nb = (n + T - 1) / T;
    uint32_t h_counters[3];
    uint32_t depth = 0;

    out.release();
    out.n_tris = n;
    CK(salloc(st, &wide, std::max<uint32_t>(n, 1)));
    CK(dalloc(&tri_out, n));
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
    k_prep_tris<<<nb, T, 0, st>>>(sv, tri_world, prim_lo, prim_hi, bounds);
    *launches += 2;
    if (n == 1) {
        k_single_tri_root<<<1, 1, 0, st>>>(prim_lo, prim_hi, tri_world, wide, tri_out);
        *launches += 1;
        out.n_wide = 1;
        out.depth = 1;
        goto finish;
    }
    CK(salloc(st, &keys, n));
    CK(salloc(st, &keys_tmp, n));
    CK(salloc(st, &vals, n));
    CK(salloc(st, &vals_tmp, n));
    CK(salloc(st, &tile_hist, onesweep_scratch_words(n)));
    k_morton<<<nb, T, 0, st>>>(prim_lo, prim_hi, n, bounds, keys, vals);
    *launches += 1;
    CK(radix_sort_pairs(st, keys, vals, keys_tmp, vals_tmp, n, 63, tile_hist, launches));
    CK(salloc(st, &bn, 2 * (size_t)n - 1));
    CK(salloc(st, &parent, 2 * (size_t)n - 1));
    CK(salloc(st, &node_count, 2 * (size_t)n - 1));
    CK(salloc(st, &node_cost, 2 * (size_t)n - 1));
    CK(salloc(st, &flags, n));
    k_hierarchy<<<nb, T, 0, st>>>(keys, vals, prim_lo, prim_hi, (int)n, bn, parent, node_count, node_cost, flags);
    k_bottom_up<<<nb, T, 0, st>>>((int)n, bn, parent, node_count, node_cost, flags, 0, 0);
    *launches += 2;
    CK(cudaMemcpyAsync(&out.sah_lbvh, node_cost, sizeof(float), cudaMemcpyDeviceToHost, st));
    for (int pass = 0; pass < opt.treelet_passes; pass++) {
        k_clear_u32<<<nb, T, 0, st>>>(flags, n);
        if (opt.coop_treelet)
            k_bottom_up_coop<<<(n + TL_BLOCK - 1) / TL_BLOCK, TL_BLOCK, 0, st>>>((int)n, bn, parent, node_count, node_cost, flags, opt.treelet_gamma);
        else
            k_bottom_up<<<(n + 63) / 64, 64, 0, st>>>((int)n, bn, parent, node_count, node_cost, flags, 1, opt.treelet_gamma);
        *launches += 2;
    }
    CK(cudaMemcpyAsync(&out.sah_final, node_cost, sizeof(float), cudaMemcpyDeviceToHost, st));
    {
        BNode root;
        CK(cudaMemcpyAsync(&root, bn, sizeof(BNode), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const float a = half_area(root.lo, root.hi);
        if (a > 0.0f) { out.sah_lbvh /= a; out.sah_final /= a; }
        out.lo[0] = root.lo.x; out.lo[1] = root.lo.y; out.lo[2] = root.lo.z;
        out.hi[0] = root.hi.x; out.hi[1] = root.hi.y; out.hi[2] = root.hi.z;
    }
    // collapse, one launch per level of the wide tree
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
        uint32_t n_items = 1;
        while (n_items) {
            k_collapse_level<<<(n_items + 63) / 64, 64, 0, st>>>(bn, node_count, (int)n - 1, queue_a, n_items, wide, counters, vals,
                                                                 tri_world, tri_out, queue_b);
            *launches += 1;
            CK(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            n_items = h_counters[2];
            h_counters[2] = 0;
            CK(cudaMemcpyAsync(counters + 2, &h_counters[2], sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            std::swap(queue_a, queue_b);
            depth++;
            if (depth > SOLB_MAX_WIDE_DEPTH) { err = cudaErrorLaunchOutOfResources; goto done; }
        }
        out.n_wide = h_counters[0];
        out.depth = depth;
        if (h_counters[1] != n) { err = cudaErrorUnknown; goto done; }  // every triangle must land in exactly one leaf
    }
finish:
    CK(cudaStreamSynchronize(st));
    // shrink the node array to its final size
    {
        Node8 *final_nodes = nullptr;
        CK(dalloc(&final_nodes, out.n_wide));
        CK(cudaMemcpyAsync(final_nodes, wide, sizeof(Node8) * out.n_wide, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        out.nodes = final_nodes;
        out.tris = tri_out;
        tri_out = nullptr;
        out.n_binary = n >= 2 ? 2 * n - 1 : n;
    }
done:
    {
        void *scratch[] = { tri_world, prim_lo, prim_hi, bounds, keys, keys_tmp, vals, vals_tmp, tile_hist, flags, counters, bn,
                            parent, node_count, node_cost, queue_a, queue_b, wide };
        for (void *q : scratch)
            if (q) cudaFreeAsync(q, st);
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
