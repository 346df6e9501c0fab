// build.cuh — per-element logic of the GPU BVH builder (Morton code, Karras hierarchy, treelet
// restructuring, 8-wide collapse).  Replaces the driver builds behind BLAS::new / TLAS::new /
// TLAS::regenerate (src/ray/acceleration.rs:136-239,344-467 -> vkCmdBuildAccelerationStructuresKHR).
//
// The reference creates one BLAS and exactly one instance per primitive section
// (src/ray/mod.rs:122 "TODO: support multiple instances per BLAS"), so instancing saves nothing;
// the B200 build therefore bakes every instance transform into world-space triangles and builds
// ONE hierarchy over all of them ("flattened TLAS"): no per-instance ray transform, no overlap
// penalty between instances whose boxes intersect (tunnel.gltf's two sections share one box).
//
// Functions here are SOLB_HD: kernels in build.cu call them per thread; tests/emu runs the same
// code sequentially on the CPU.
#pragma once
#include "bvh.cuh"

namespace solb {

// ---- binary hierarchy (Karras 2012 numbering: internal i in [0, n-2], leaf j -> n-1+j) ----------
struct BNode {
    float3 lo; int left;   // child node ids (unified numbering); leaves: left = right = -1
    float3 hi; int right;
};
static_assert(sizeof(BNode) == 32, "BNode must be 32 bytes");

SOLB_HD uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// 63-bit Morton code of a centroid inside [lo, hi]
SOLB_HD uint64_t morton63(float3 c, float3 lo, float3 inv_ext) {
    const float scale = 2097152.0f;  // 2^21
    float fx = (c.x - lo.x) * inv_ext.x * scale, fy = (c.y - lo.y) * inv_ext.y * scale, fz = (c.z - lo.z) * inv_ext.z * scale;
    uint32_t x = (uint32_t)fminf(fmaxf(fx, 0.0f), scale - 1.0f);
    uint32_t y = (uint32_t)fminf(fmaxf(fy, 0.0f), scale - 1.0f);
    uint32_t z = (uint32_t)fminf(fmaxf(fz, 0.0f), scale - 1.0f);
    return (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
}

// length of the common prefix of keys i and j, index-augmented so duplicate keys still split
SOLB_HD int karras_delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz64(a ^ b);
}

// internal node i of n-1: children and covered range [first, last] of sorted primitives
SOLB_HD void karras_node(const uint64_t *keys, int n, int i, int &left, int &right, int &first, int &last) {
    const int d = (karras_delta(keys, n, i, i + 1) - karras_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = karras_delta(keys, n, i, i - d);
    int lmax = 2;
    while (karras_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (karras_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = karras_delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (karras_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    first = i < j ? i : j;
    last = i < j ? j : i;
    left = (first == gamma) ? (n - 1 + gamma) : gamma;
    right = (last == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
}

// ---- treelet restructuring (Karras & Aila 2013, "Fast Parallel Construction of High-Quality
//      Bounding Volume Hierarchies"), one treelet of up to 7 leaves per call, exhaustive DP over
//      the 2^7 leaf subsets.  Internal node ids of the treelet are reused, so the rest of the tree
//      (parents, ranges are NOT preserved: callers must not rely on Karras ranges afterwards). --------
#define SOLB_TREELET_N 7
#define SOLB_SAH_CI 1.2f  // cost of an internal node visit
#define SOLB_SAH_CT 1.0f  // cost per triangle

struct TreeletScratch {
    float area[128];
    float cost[128];
    uint8_t part[128];
};

// subtree SAH cost + triangle count are kept per node by the caller (cost[], count[])
SOLB_HD float leaf_or_internal_cost(float area, float children_cost, int count) {
    // a subtree may be flattened into a leaf if it is small enough for the 8-wide collapse
    float c = SOLB_SAH_CI * area + children_cost;
    if (count <= SOLB_MAX_LEAF_TRIS) c = fminf(c, SOLB_SAH_CT * area * (float)count);
    return c;
}

// Restructure the treelet rooted at `root` (an internal node).  node_cost/node_count are per-node
// subtree SAH cost / triangle count (valid for all descendants on entry; updated for the treelet's
// internal nodes on exit).  Returns the new cost of root.
SOLB_HD float optimize_treelet(BNode *bn, int *parent, float *node_cost, int *node_count, int n_internal, int root,
                               TreeletScratch &sc) {
    int leaves[SOLB_TREELET_N];
    int internals[SOLB_TREELET_N - 1];
    int nl = 0, ni = 0;
    internals[ni++] = root;
    leaves[nl++] = bn[root].left;
    leaves[nl++] = bn[root].right;
    while (nl < SOLB_TREELET_N) {
        int best = -1;
        float best_a = -1.0f;
        for (int i = 0; i < nl; i++) {
            const int c = leaves[i];
            if (c < n_internal) {
                const float a = half_area(bn[c].lo, bn[c].hi);
                if (a > best_a) { best_a = a; best = i; }
            }
        }
        if (best < 0) break;
        const int c = leaves[best];
        internals[ni++] = c;
        leaves[best] = bn[c].left;
        leaves[nl++] = bn[c].right;
    }
    if (nl < 3) return node_cost[root];  // nothing to restructure
    const int full = (1 << nl) - 1;
    // areas of every subset's union box
    for (int s = 1; s <= full; s++) {
        float3 lo = f3(3.4e38f, 3.4e38f, 3.4e38f), hi = f3(-3.4e38f, -3.4e38f, -3.4e38f);
        for (int i = 0; i < nl; i++)
            if (s & (1 << i)) { lo = fmin3(lo, bn[leaves[i]].lo); hi = fmax3(hi, bn[leaves[i]].hi); }
        sc.area[s] = half_area(lo, hi);
    }
    // singletons
    for (int i = 0; i < nl; i++) { sc.cost[1 << i] = node_cost[leaves[i]]; sc.part[1 << i] = 0; }
    // subsets by increasing popcount
    for (int k = 2; k <= nl; k++) {
        for (int s = 1; s <= full; s++) {
            if (popc32((uint32_t)s) != k) continue;
            float best_c = 3.4e38f;
            int best_p = 0;
            // enumerate partitions (p, s^p) with the lowest set bit fixed in p to halve the work
            const int delta = (s - 1) & s;
            int p = (-delta) & s;
            do {
                const float c = sc.cost[p] + sc.cost[s ^ p];
                if (c < best_c) { best_c = c; best_p = p; }
                p = (p - delta) & s;
            } while (p != 0);
            int cnt = 0;
            for (int i = 0; i < nl; i++) if (s & (1 << i)) cnt += node_count[leaves[i]];
            sc.cost[s] = leaf_or_internal_cost(sc.area[s], best_c, cnt);
            sc.part[s] = (uint8_t)best_p;
        }
    }
    const float old_cost = node_cost[root];
    if (!(sc.cost[full] < old_cost * 0.9999f)) return old_cost;  // keep the topology unless it improves
    // rebuild topology top-down reusing the internal node ids
    int stack_set[SOLB_TREELET_N], stack_node[SOLB_TREELET_N];
    int sp = 0, next_internal = 1;
    stack_set[sp] = full; stack_node[sp] = root; sp++;
    // process in an order that lets us fix boxes afterwards: record (node) in creation order
    int order[SOLB_TREELET_N - 1];
    int n_order = 0;
    while (sp) {
        sp--;
        const int s = stack_set[sp], node = stack_node[sp];
        order[n_order++] = node;
        const int p0 = sc.part[s], p1 = s ^ p0;
        int child[2];
        const int sub[2] = { p0, p1 };
        for (int k = 0; k < 2; k++) {
            if (popc32((uint32_t)sub[k]) == 1) {
                child[k] = leaves[bfind32((uint32_t)sub[k])];
            } else {
                child[k] = internals[next_internal++];
                stack_set[sp] = sub[k]; stack_node[sp] = child[k]; sp++;
            }
            parent[child[k]] = node;
        }
        bn[node].left = child[0];
        bn[node].right = child[1];
    }
    // boxes / costs / counts bottom-up (reverse creation order: children are created after parents)
    for (int k = n_order - 1; k >= 0; k--) {
        const int node = order[k];
        const int l = bn[node].left, r = bn[node].right;
        bn[node].lo = fmin3(bn[l].lo, bn[r].lo);
        bn[node].hi = fmax3(bn[l].hi, bn[r].hi);
        node_count[node] = node_count[l] + node_count[r];
        node_cost[node] = leaf_or_internal_cost(half_area(bn[node].lo, bn[node].hi), node_cost[l] + node_cost[r], node_count[node]);
    }
    return node_cost[root];
}

// ---- optimal wide collapse (Ylitie, Karras, Laine 2017, section 4.1): dynamic program over the binary tree ----
// c(n, i) = cheapest way to represent the subtree of binary node n as at most i children of a wide node:
//   c(n, 1) = min(leaf cost, CN * area(n) + distribute(n, 8))      n becomes ONE child: a leaf or an 8-wide node of its own
//   c(n, i) = min(distribute(n, i), c(n, i - 1))                    i = 2..7: n's two subtrees share i child slots
//   distribute(n, j) = min over 0 < k < j of c(left, k) + c(right, j - k)
// Cost unit: one triangle test of a unit-area box (CP = 1); CN = an 8-wide node visit (the node step executes ~2x the
// instructions of a triangle step in k_wf_trace).  Measured on tunnel.gltf bounce rays against the greedy
// largest-area-first collapse: 30 % fewer wide nodes, 5 % fewer node visits per ray.
#define SOLB_DP_CN 2.0f
#define SOLB_DP_CP 1.0f
struct DpEntry {
    float c[7];   // c[i - 1] = c(n, i)
    uint32_t k;   // bits [3 (j - 2), 3 (j - 2) + 2], j = 2..8: the split k of distribute(n, j); 0 = "same as j - 1"; bit 31: n is a leaf
};
static_assert(sizeof(DpEntry) == 32, "DpEntry must be 32 bytes");

// cost model of one tree: triangles (CP = 1, leaves of up to 3) or, for a TLAS, instances: entering an instance costs a whole
// BLAS walk, so its CP is much larger and every instance gets a leaf child (and a box test) of its own
struct DpCost {
    float cn, cp;
    int max_leaf;
};
SOLB_HD DpCost dp_cost_triangles() { DpCost c; c.cn = SOLB_DP_CN; c.cp = SOLB_DP_CP; c.max_leaf = SOLB_MAX_LEAF_TRIS; return c; }

SOLB_HD void dp_leaf_entry(DpEntry &e, float area, DpCost cost = dp_cost_triangles()) {
    for (int i = 0; i < 7; i++) e.c[i] = cost.cp * area;
    e.k = 0x80000000u;
}

SOLB_HD void dp_inner_entry(DpEntry &e, const DpEntry &l, const DpEntry &r, float area, int count, DpCost cost = dp_cost_triangles()) {
    float dist[9];
    uint32_t kbits = 0;
    for (int j = 2; j <= 8; j++) {
        float best = 3.4e38f;
        int bk = 1;
        for (int k = 1; k < j; k++) {
            const int a = k > 7 ? 7 : k, b = (j - k) > 7 ? 7 : (j - k);
            const float c = l.c[a - 1] + r.c[b - 1];
            if (c < best) { best = c; bk = k; }
        }
        dist[j] = best;
        kbits |= (uint32_t)bk << (3 * (j - 2));
    }
    const float c_leaf = count <= cost.max_leaf ? cost.cp * area * (float)count : 3.4e38f;
    const float c_node = dist[8] + cost.cn * area;
    if (c_leaf <= c_node) kbits |= 0x80000000u;
    e.c[0] = fminf(c_leaf, c_node);
    for (int i = 2; i <= 7; i++) {
        if (dist[i] < e.c[i - 2]) e.c[i - 1] = dist[i];
        else { e.c[i - 1] = e.c[i - 2]; kbits &= ~(7u << (3 * (i - 2))); }
    }
    e.k = kbits;
}

// children of the wide node rooted at binary node `root` according to the DP decisions; returns their count (2..8)
SOLB_HD int dp_gather_children(const BNode *bn, const DpEntry *dp, int n_internal, int root, int *cand) {
    int st_node[8], st_budget[8];
    int sp = 0, n = 0;
    {
        const int k = (int)((dp[root].k >> 18) & 7u);  // distribute(root, 8)
        st_node[sp] = bn[root].right; st_budget[sp] = 8 - k; sp++;
        st_node[sp] = bn[root].left; st_budget[sp] = k; sp++;
    }
    while (sp) {
        sp--;
        const int node = st_node[sp];
        int b = st_budget[sp] > 7 ? 7 : st_budget[sp];
        if (node >= n_internal) { cand[n++] = node; continue; }
        const uint32_t kb = dp[node].k;
        while (b > 1 && ((kb >> (3 * (b - 2))) & 7u) == 0u) b--;
        if (b == 1) { cand[n++] = node; continue; }
        const int k = (int)((kb >> (3 * (b - 2))) & 7u);
        st_node[sp] = bn[node].right; st_budget[sp] = b - k; sp++;
        st_node[sp] = bn[node].left; st_budget[sp] = k; sp++;
    }
    return n;
}

// ---- collapse: binary tree -> 8-wide compressed nodes ------------------------------------------------
#if defined(__CUDA_ARCH__)
SOLB_HD uint32_t counter_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
#else
SOLB_HD uint32_t counter_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
#endif

struct CollapseItem {
    int bnode;       // binary node to turn into a wide node (always an internal binary node)
    uint32_t wnode;  // index of the wide node to write
};

// Gather the triangles under binary node c (in-order) into out[], return count (<= SOLB_MAX_LEAF_TRIS)
SOLB_HD int gather_leaf_tris(const BNode *bn, int n_internal, int c, int *out) {
    int n = 0;
    int stack[8];
    int sp = 0;
    stack[sp++] = c;
    while (sp) {
        const int x = stack[--sp];
        if (x >= n_internal) out[n++] = x - n_internal;  // sorted-position of the leaf primitive
        else { stack[sp++] = bn[x].right; stack[sp++] = bn[x].left; }
    }
    return n;
}

// Slot auction of the wide node's children (after Ylitie et al. 2017): repeatedly give the (child, free slot) pair with the
// largest dot(child centre - node centre, octant direction of the slot) its slot, so children are visited roughly front to
// back when the slots are walked in the ray's octant order.  Written with fixed-bound loops and bit masks so the 8 x 8 scores
// stay in registers: one thread per wide node runs this on the critical path of every collapse level.
SOLB_HD void assign_child_slots(const float3 *rel, int n, int *slot_of) {
    uint32_t free_child = (1u << n) - 1u, free_slot = 0xffu;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) slot_of[i] = -1;
    for (int round = 0; round < n; round++) {
        float best_v = -3.4e38f;
        int bi = -1, bs = -1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 8; i++) {
            if (!((free_child >> i) & 1u)) continue;
            const float rx = rel[i].x, ry = rel[i].y, rz = rel[i].z;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int sl = 0; sl < 8; sl++) {
                if (!((free_slot >> sl) & 1u)) continue;
                const float v = ((sl & 4) ? rx : -rx) + ((sl & 2) ? ry : -ry) + ((sl & 1) ? rz : -rz);
                if (v > best_v) { best_v = v; bi = i; bs = sl; }
            }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 8; i++)
            if (i == bi) slot_of[i] = bs;
        free_child &= ~(1u << bi);
        free_slot &= ~(1u << bs);
    }
}

// One work item.  sorted_prim[j] = global triangle id at sorted position j; tri_world indexed by
// global triangle id; tri_out in wide-leaf order.  TLAS builds pass leaf_prim_out instead of tri_world / tri_out:
// the primitive (instance) id of every leaf slot, from which the caller writes its own leaf records.
SOLB_HD void collapse_one(const BNode *bn, const int *node_count, int n_internal, CollapseItem item, Node8 *wide,
                          uint32_t *wide_count, uint32_t *tri_count, const uint32_t *sorted_prim, const Tri48 *tri_world,
                          Tri48 *tri_out, CollapseItem *queue_out, uint32_t *queue_out_count, uint32_t *leaf_prim_out = nullptr,
                          const DpEntry *dp = nullptr) {
    int cand[8];
    int n = 2;
    const BNode self = bn[item.bnode];
    cand[0] = self.left;
    cand[1] = self.right;
    if (dp) n = dp_gather_children(bn, dp, n_internal, item.bnode, cand);
    else {  // greedy: open the internal candidate with the largest box until 8 children
        float area[8];  // < 0: a binary leaf, cannot be opened
        for (int i = 0; i < 2; i++) area[i] = cand[i] < n_internal ? half_area(bn[cand[i]].lo, bn[cand[i]].hi) : -1.0f;
        while (n < 8) {
            int best = -1;
            float best_a = -1.0f;
            for (int i = 0; i < n; i++)
                if (area[i] > best_a) { best_a = area[i]; best = i; }
            if (best < 0) break;
            const int c = cand[best], l = bn[c].left, r = bn[c].right;
            cand[best] = l;
            area[best] = l < n_internal ? half_area(bn[l].lo, bn[l].hi) : -1.0f;
            cand[n] = r;
            area[n] = r < n_internal ? half_area(bn[r].lo, bn[r].hi) : -1.0f;
            n++;
        }
    }
    // children: boxes, leaf / inner decision, slots
    const float3 nlo = self.lo, nhi = self.hi;
    const float3 nc = (nlo + nhi) * 0.5f;
    float3 clo[8], chi[8], rel[8];
    int ccount[8];
    bool cleaf[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) {
        if (i < n) {
            const int c = cand[i];
            const BNode b = bn[c];
            clo[i] = b.lo; chi[i] = b.hi;
            rel[i] = (b.lo + b.hi) * 0.5f - nc;
            ccount[i] = node_count[c];
            cleaf[i] = dp ? (dp[c].k >> 31) != 0u : ccount[i] <= SOLB_MAX_LEAF_TRIS;
        } else {
            clo[i] = chi[i] = rel[i] = f3(0.0f, 0.0f, 0.0f);
            ccount[i] = 0;
            cleaf[i] = false;
        }
    }
    int slot_of[8];
    assign_child_slots(rel, n, slot_of);
    ChildRef ch[8];
    int child_of_slot[8];
    uint32_t n_inner = 0, n_tris = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int sl = 0; sl < 8; sl++) {
        ch[sl].valid = 0;
        child_of_slot[sl] = -1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 8; i++) {
            if (i < n && slot_of[i] == sl) {
                child_of_slot[sl] = cand[i];
                ch[sl].valid = 1;
                ch[sl].lo = clo[i];
                ch[sl].hi = chi[i];
                if (cleaf[i]) {
                    ch[sl].is_inner = 0; ch[sl].tri_offset = n_tris; ch[sl].tri_count = (uint32_t)ccount[i];
                    n_tris += (uint32_t)ccount[i];
                } else {
                    ch[sl].is_inner = 1; ch[sl].tri_offset = 0; ch[sl].tri_count = 0;
                    n_inner++;
                }
            }
        }
    }
    const uint32_t child_base = n_inner ? counter_add(wide_count, n_inner) : 0u;
    const uint32_t tri_base = n_tris ? counter_add(tri_count, n_tris) : 0u;
    encode_node8(wide[item.wnode], nlo, nhi, child_base, tri_base, ch);
    uint32_t q = n_inner ? counter_add(queue_out_count, n_inner) : 0u;
    uint32_t k = 0;
    for (int sl = 0; sl < 8; sl++) {
        const int c = child_of_slot[sl];
        if (c < 0) continue;
        if (ch[sl].is_inner) {
            CollapseItem it;
            it.bnode = c;
            it.wnode = child_base + k;
            queue_out[q + k] = it;
            k++;
        } else {
            int prims[SOLB_MAX_LEAF_TRIS + 1];
            const int m = gather_leaf_tris(bn, n_internal, c, prims);
            for (int j = 0; j < m; j++) {
                const uint32_t dst = tri_base + ch[sl].tri_offset + (uint32_t)j, prim = sorted_prim[prims[j]];
                if (leaf_prim_out) leaf_prim_out[dst] = prim;
                else tri_out[dst] = tri_world[prim];
            }
        }
    }
}

// World-space box of an object-space box under a column-major 4x4 (instance AABB for the TLAS)
SOLB_HD void transform_box(const float *m, float3 lo, float3 hi, float3 &out_lo, float3 &out_hi) {
    out_lo = f3(3.4e38f, 3.4e38f, 3.4e38f);
    out_hi = f3(-3.4e38f, -3.4e38f, -3.4e38f);
    for (int c = 0; c < 8; c++) {
        const float3 p = mat4_mul_point(m, f3((c & 1) ? hi.x : lo.x, (c & 2) ? hi.y : lo.y, (c & 4) ? hi.z : lo.z));
        out_lo = fmin3(out_lo, p);
        out_hi = fmax3(out_hi, p);
    }
}

// InstLeaf of instance `id`: rows of inverse(transform) (world -> object) + BLAS root + id.
// inverse() by cofactors of the upper 3x3 (affine instance matrices only, like VkAccelerationStructureInstanceKHR's 3x4).
SOLB_HD InstLeaf make_inst_leaf(const float *m, uint32_t blas_root, uint32_t id) {
    const float a = m[0], b = m[4], c = m[8], d = m[1], e = m[5], f = m[9], g = m[2], h = m[6], i = m[10];
    const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const float det = a * A + b * B + c * C;
    const float r = det != 0.0f ? 1.0f / det : 0.0f;
    // inverse 3x3, row-major
    const float i00 = A * r, i01 = -(b * i - c * h) * r, i02 = (b * f - c * e) * r;
    const float i10 = B * r, i11 = (a * i - c * g) * r, i12 = -(a * f - c * d) * r;
    const float i20 = C * r, i21 = -(a * h - b * g) * r, i22 = (a * e - b * d) * r;
    const float tx = m[12], ty = m[13], tz = m[14];
    InstLeaf L;
    L.r0 = make_float4(i00, i01, i02, -(i00 * tx + i01 * ty + i02 * tz));
    L.r1 = make_float4(i10, i11, i12, -(i10 * tx + i11 * ty + i12 * tz));
    L.r2 = make_float4(i20, i21, i22, -(i20 * tx + i21 * ty + i22 * tz));
    L.r3 = make_float4(u2f(blas_root), u2f(id), 0.0f, 0.0f);
    return L;
}

}  // namespace solb
