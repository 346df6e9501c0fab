// emu.cpp — TEST-ONLY host emulation of libsolb's per-thread device logic.
// Compiles sol_rs_b200/csrc/{bvh,build,shade}.cuh with g++ (their functions are SOLB_HD) and steps
// them sequentially so builder / traversal / shading logic can be checked against the oracle in the
// CPU-only test tier.  It is NOT a product path: the library itself has no CPU fallback and nothing in
// sol_rs_b200/ links or loads this file.  Kernel-side plumbing (queues, atomics, radix sort) is only
// testable on the GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "build.cuh"
#include "shade.cuh"

using namespace solb;

struct EmuStack {
    uint2 e[128];
    int sp = 0;
    int max_sp = 0;
    void push(uint2 v) { e[sp++] = v; if (sp > max_sp) max_sp = sp; }
    uint2 pop() { return e[--sp]; }
    bool empty() const { return sp == 0; }
};

struct EmuScene {
    std::vector<DeviceInstance> inst;
    std::vector<DeviceBlas> blas;
    std::vector<uint32_t> first_tri;  // per-instance triangle prefix (flattened build)
    std::vector<float4> vertices;     // 4 per vertex
    std::vector<uint32_t> indices;
    std::vector<ShadeRecord> shade;   // per geometry triangle
    std::vector<Node8> nodes;
    std::vector<Tri48> tris;
    std::vector<InstLeaf> inst_leaves;
    std::vector<std::vector<float4>> texels;  // base-colour textures (linear light), as solb_scene_set_textures uploads them
    std::vector<TexDesc> tex;
    bool two_level = false;
    uint32_t n_tris = 0, n_geom_tris = 0, depth = 0, tlas_depth = 0;
    float sah_lbvh = 0, sah_final = 0;
    int max_stack = 0;
};

// binary radix tree over n >= 2 boxes: Morton sort, Karras links, refit, treelet passes (same per-element code as build.cu)
struct EmuTree {
    std::vector<BNode> bn;
    std::vector<int> parent, count;
    std::vector<float> cost;
    std::vector<uint32_t> vals;  // sorted position -> primitive
    std::vector<DpEntry> dp;     // optimal-collapse table (empty: greedy collapse)
    int ni = 0;
    float sah_lbvh = 0, sah_final = 0;
};

static int g_emu_dp = 1;  // mirrors BuildOptions::dp_collapse

static void emu_binary_tree(const std::vector<float3> &plo, const std::vector<float3> &phi, int treelet_passes, int gamma, EmuTree &T,
                            DpCost cost = dp_cost_triangles()) {
    const uint32_t n = (uint32_t)plo.size();
    float3 clo = f3(3.4e38f, 3.4e38f, 3.4e38f), chi = f3(-3.4e38f, -3.4e38f, -3.4e38f);
    for (uint32_t g = 0; g < n; g++) {
        const float3 c = (plo[g] + phi[g]) * 0.5f;
        clo = fmin3(clo, c); chi = fmax3(chi, c);
    }
    const float3 ext = chi - clo;
    const float3 inv = f3(ext.x > 0 ? 1.0f / ext.x : 0, ext.y > 0 ? 1.0f / ext.y : 0, ext.z > 0 ? 1.0f / ext.z : 0);
    std::vector<uint64_t> keys(n);
    for (uint32_t g = 0; g < n; g++) keys[g] = morton63((plo[g] + phi[g]) * 0.5f, clo, inv);
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> skeys(n);
    T.vals.resize(n);
    for (uint32_t i = 0; i < n; i++) { skeys[i] = keys[order[i]]; T.vals[i] = order[i]; }
    const int ni = (int)n - 1;
    T.ni = ni;
    T.bn.assign(2 * n - 1, BNode());
    T.parent.assign(2 * n - 1, -1);
    T.count.assign(2 * n - 1, 0);
    T.cost.assign(2 * n - 1, 0.0f);
    auto &bn = T.bn;
    for (uint32_t i = 0; i < n; i++) {
        BNode &l = bn[ni + i];
        l.lo = plo[T.vals[i]]; l.hi = phi[T.vals[i]]; l.left = l.right = -1;
        T.count[ni + i] = 1;
        T.cost[ni + i] = SOLB_SAH_CT * half_area(l.lo, l.hi);
    }
    for (int i = 0; i < ni; i++) {
        int l, r, first, last;
        karras_node(skeys.data(), (int)n, i, l, r, first, last);
        bn[i].left = l; bn[i].right = r;
        T.parent[l] = i; T.parent[r] = i;
    }
    // bottom-up order = reverse BFS from the root
    auto bottom_up = [&](int mode) {
        std::vector<int> bfs;
        bfs.push_back(0);
        for (size_t k = 0; k < bfs.size(); k++) {
            int x = bfs[k];
            if (x < ni) { bfs.push_back(bn[x].left); bfs.push_back(bn[x].right); }
        }
        TreeletScratch sc;
        for (size_t k = bfs.size(); k-- > 0;) {
            int node = bfs[k];
            if (node >= ni) continue;
            int l = bn[node].left, r = bn[node].right;
            bn[node].lo = fmin3(bn[l].lo, bn[r].lo);
            bn[node].hi = fmax3(bn[l].hi, bn[r].hi);
            T.count[node] = T.count[l] + T.count[r];
            T.cost[node] = leaf_or_internal_cost(half_area(bn[node].lo, bn[node].hi), T.cost[l] + T.cost[r], T.count[node]);
            if (mode == 1 && T.count[node] >= gamma) optimize_treelet(bn.data(), T.parent.data(), T.cost.data(), T.count.data(), ni, node, sc);
        }
    };
    bottom_up(0);
    const float root_area = half_area(bn[0].lo, bn[0].hi);
    T.sah_lbvh = root_area > 0 ? T.cost[0] / root_area : 0;
    for (int p = 0; p < treelet_passes; p++) bottom_up(1);
    T.sah_final = root_area > 0 ? T.cost[0] / root_area : 0;
    if (g_emu_dp) {  // k_bottom_up_dp
        T.dp.assign(2 * n - 1, DpEntry());
        std::vector<int> bfs;
        bfs.push_back(0);
        for (size_t k = 0; k < bfs.size(); k++) {
            int x = bfs[k];
            if (x < ni) { bfs.push_back(bn[x].left); bfs.push_back(bn[x].right); }
        }
        for (size_t k = bfs.size(); k-- > 0;) {
            const int node = bfs[k];
            const float a = half_area(bn[node].lo, bn[node].hi);
            if (node >= ni) dp_leaf_entry(T.dp[node], a, cost);
            else dp_inner_entry(T.dp[node], T.dp[bn[node].left], T.dp[bn[node].right], a, T.count[node], cost);
        }
    }
}

// level-synchronous collapse of T's root into wide[root_wnode]; returns the number of levels
static uint32_t emu_collapse(const EmuTree &T, uint32_t root_wnode, Node8 *wide, uint32_t *wide_count, uint32_t *tri_count,
                             const Tri48 *src, Tri48 *dst, uint32_t *leaf_prim) {
    const uint32_t n = (uint32_t)T.ni + 1;
    std::vector<CollapseItem> qa(n), qb(n);
    uint32_t nq = 1, depth = 0;
    qa[0].bnode = 0; qa[0].wnode = root_wnode;
    while (nq) {
        uint32_t nout = 0;
        for (uint32_t i = 0; i < nq; i++)
            collapse_one(T.bn.data(), T.count.data(), T.ni, qa[i], wide, wide_count, tri_count, T.vals.data(), src, dst, qb.data(), &nout,
                         leaf_prim, T.dp.empty() ? nullptr : T.dp.data());
        std::swap(qa, qb);
        nq = nout;
        depth++;
    }
    return depth;
}

static void single_leaf_root(Node8 &out, float3 lo, float3 hi, uint32_t tri_base) {
    ChildRef ch[8];
    for (auto &c : ch) c.valid = 0;
    ch[0].valid = 1; ch[0].lo = lo; ch[0].hi = hi; ch[0].is_inner = 0; ch[0].tri_offset = 0; ch[0].tri_count = 1;
    encode_node8(out, lo, hi, 0, tri_base, ch);
}

extern "C" {

// instances carry their BLAS id in .blas (ids must appear in increasing order of first use); geometry ranges are per instance
EmuScene *emu_scene_create(uint32_t n_inst, const DeviceInstance *inst, const float *vertices, uint32_t n_vertices,
                           const uint32_t *indices, uint32_t n_indices) {
    EmuScene *s = new EmuScene();
    s->inst.assign(inst, inst + n_inst);
    s->vertices.resize((size_t)n_vertices * 4);
    memcpy(s->vertices.data(), vertices, (size_t)n_vertices * 64);
    s->indices.assign(indices, indices + n_indices);
    s->first_tri.push_back(0);
    for (uint32_t i = 0; i < n_inst; i++) {
        DeviceInstance &di = s->inst[i];
        if (di.blas >= s->blas.size()) {
            DeviceBlas db;
            db.first_vertex = di.first_vertex; db.first_index = di.first_index; db.n_indices = di.n_indices; db.first_tri = s->n_geom_tris;
            di.blas = (uint32_t)s->blas.size();
            s->blas.push_back(db);
            s->n_geom_tris += di.n_indices / 3;
        }
        di.shade_first_tri = s->blas[di.blas].first_tri;
        s->first_tri.push_back(s->first_tri.back() + di.n_indices / 3);
    }
    s->n_tris = s->first_tri.back();
    s->shade.resize(std::max<uint32_t>(s->n_geom_tris, 1));
    for (const DeviceBlas &db : s->blas)
        for (uint32_t p = 0; p < db.n_indices / 3; p++) {
            float f[28];
            for (int k = 0; k < 3; k++) {
                uint32_t vi = db.first_vertex + indices[db.first_index + 3 * p + k];
                const float4 pos = s->vertices[4 * (size_t)vi], col = s->vertices[4 * (size_t)vi + 1], nrm = s->vertices[4 * (size_t)vi + 2];
                f[9 * k] = pos.x; f[9 * k + 1] = pos.y; f[9 * k + 2] = pos.z;
                f[9 * k + 3] = nrm.x; f[9 * k + 4] = nrm.y; f[9 * k + 5] = nrm.z;
                f[9 * k + 6] = col.x; f[9 * k + 7] = col.y; f[9 * k + 8] = col.z;
            }
            f[27] = 0;
            ShadeRecord &r = s->shade[db.first_tri + p];
            for (int q = 0; q < 7; q++) r.q[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        }
    return s;
}

void emu_scene_destroy(EmuScene *s) { delete s; }
void emu_set_dp_collapse(int on) { g_emu_dp = on; }

void emu_set_transform(EmuScene *s, uint32_t i, const float *t, const float *t_it) {
    memcpy(s->inst[i].transform, t, 64);
    memcpy(s->inst[i].transform_it, t_it, 64);
}

// Same pipeline as build.cu (flattened mode), run sequentially.  treelet_passes: 0 = plain LBVH.
int emu_build(EmuScene *s, int treelet_passes, int gamma) {
    const uint32_t n = s->n_tris;
    s->nodes.clear(); s->tris.clear(); s->inst_leaves.clear();
    s->two_level = false;
    if (n == 0) {
        ChildRef ch[8];
        for (auto &c : ch) c.valid = 0;
        s->nodes.resize(1);
        encode_node8(s->nodes[0], f3(0, 0, 0), f3(0, 0, 0), 0, 0, ch);
        s->depth = 1;
        return 0;
    }
    std::vector<Tri48> tri_world(n);
    std::vector<float3> plo(n), phi(n);
    for (uint32_t i = 0; i < s->inst.size(); i++)
        for (uint32_t p = 0; p < s->inst[i].n_indices / 3; p++) {
            const uint32_t g = s->first_tri[i] + p;
            float3 v[3];
            for (int k = 0; k < 3; k++) {
                uint32_t vi = s->inst[i].first_vertex + s->indices[s->inst[i].first_index + 3 * p + k];
                const float4 pos = s->vertices[4 * (size_t)vi];
                v[k] = mat4_mul_point(s->inst[i].transform, f3(pos.x, pos.y, pos.z));
            }
            tri_world[g].v0 = make_float4(v[0].x, v[0].y, v[0].z, u2f(i));
            tri_world[g].v1 = make_float4(v[1].x, v[1].y, v[1].z, u2f(p));
            tri_world[g].v2 = make_float4(v[2].x, v[2].y, v[2].z, u2f(s->inst[i].shade_first_tri + p));
            plo[g] = fmin3(v[0], fmin3(v[1], v[2]));
            phi[g] = fmax3(v[0], fmax3(v[1], v[2]));
        }
    s->tris.resize(n);
    if (n == 1) {
        s->nodes.resize(1);
        single_leaf_root(s->nodes[0], plo[0], phi[0], 0);
        s->tris[0] = tri_world[0];
        s->depth = 1;
        return 0;
    }
    EmuTree T;
    emu_binary_tree(plo, phi, treelet_passes, gamma, T);
    s->sah_lbvh = T.sah_lbvh;
    s->sah_final = T.sah_final;
    s->nodes.resize(n);
    uint32_t wide_count = 1, tri_count = 0;
    s->depth = emu_collapse(T, 0, s->nodes.data(), &wide_count, &tri_count, tri_world.data(), s->tris.data(), nullptr);
    s->nodes.resize(wide_count);
    return tri_count == n ? 0 : -1;
}

// TLAS over the instances' world boxes (build.cu: rebuild_tlas): nodes[0, tlas_cap), inst_leaves
static int emu_rebuild_tlas(EmuScene *s, const std::vector<float3> &blo, const std::vector<float3> &bhi) {
    const uint32_t n = (uint32_t)s->inst.size(), tlas_cap = std::max<uint32_t>(n, 1);
    s->inst_leaves.assign(std::max<uint32_t>(n, 1), InstLeaf());
    if (n == 0) {
        ChildRef ch[8];
        for (auto &c : ch) c.valid = 0;
        encode_node8(s->nodes[0], f3(0, 0, 0), f3(0, 0, 0), 0, 0, ch);
        s->tlas_depth = 1;
        return 0;
    }
    std::vector<float3> plo(n), phi(n);
    for (uint32_t i = 0; i < n; i++) transform_box(s->inst[i].transform, blo[s->inst[i].blas], bhi[s->inst[i].blas], plo[i], phi[i]);
    if (n == 1) {
        single_leaf_root(s->nodes[0], plo[0], phi[0], 0);
        s->inst_leaves[0] = make_inst_leaf(s->inst[0].transform, tlas_cap + s->inst[0].blas, 0);
        s->tlas_depth = 1;
        return 0;
    }
    EmuTree T;
    DpCost tlas_cost;  // build.cu rebuild_tlas: an instance entry is a whole BLAS walk, every instance gets its own leaf child
    tlas_cost.cn = 2.0f; tlas_cost.cp = 8.0f; tlas_cost.max_leaf = 1;
    emu_binary_tree(plo, phi, 2, 7, T, tlas_cost);
    s->sah_lbvh = T.sah_lbvh;
    s->sah_final = T.sah_final;
    std::vector<uint32_t> leaf_prim(n);
    uint32_t wide_count = 1, leaf_count = 0;
    s->tlas_depth = emu_collapse(T, 0, s->nodes.data(), &wide_count, &leaf_count, nullptr, nullptr, leaf_prim.data());
    if (leaf_count != n || wide_count > tlas_cap) return -1;
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t i = leaf_prim[j];
        s->inst_leaves[j] = make_inst_leaf(s->inst[i].transform, tlas_cap + s->inst[i].blas, i);
    }
    return 0;
}

// Two-level build (build.cu: build_accel_two_level): every BLAS in object space, then the TLAS.  BLASes are built one
// after another here; the GPU builds them in one batched pass with segmented Morton keys.
int emu_build_two_level(EmuScene *s, int treelet_passes, int gamma) {
    const uint32_t n = s->n_geom_tris, n_blas = (uint32_t)s->blas.size(), n_inst = (uint32_t)s->inst.size();
    const uint32_t tlas_cap = std::max<uint32_t>(n_inst, 1);
    s->two_level = true;
    s->nodes.assign((size_t)tlas_cap + n_blas + n, Node8());
    s->tris.assign(std::max<uint32_t>(n, 1), Tri48());
    uint32_t wide_count = tlas_cap + n_blas, tri_count = 0, blas_depth = 1;
    std::vector<float3> blo(n_blas), bhi(n_blas);
    for (uint32_t b = 0; b < n_blas; b++) {
        const DeviceBlas &db = s->blas[b];
        const uint32_t cnt = db.n_indices / 3;
        std::vector<Tri48> tri_obj(cnt);
        std::vector<float3> plo(cnt), phi(cnt);
        for (uint32_t p = 0; p < cnt; p++) {
            float3 v[3];
            for (int k = 0; k < 3; k++) {
                uint32_t vi = db.first_vertex + s->indices[db.first_index + 3 * p + k];
                const float4 pos = s->vertices[4 * (size_t)vi];
                v[k] = f3(pos.x, pos.y, pos.z);
            }
            tri_obj[p].v0 = make_float4(v[0].x, v[0].y, v[0].z, u2f(b));
            tri_obj[p].v1 = make_float4(v[1].x, v[1].y, v[1].z, u2f(p));
            tri_obj[p].v2 = make_float4(v[2].x, v[2].y, v[2].z, u2f(db.first_tri + p));
            plo[p] = fmin3(v[0], fmin3(v[1], v[2]));
            phi[p] = fmax3(v[0], fmax3(v[1], v[2]));
        }
        if (cnt == 1) {
            single_leaf_root(s->nodes[tlas_cap + b], plo[0], phi[0], tri_count);
            s->tris[tri_count++] = tri_obj[0];
            blo[b] = plo[0]; bhi[b] = phi[0];
            continue;
        }
        EmuTree T;
        emu_binary_tree(plo, phi, treelet_passes, gamma, T);
        blo[b] = T.bn[0].lo; bhi[b] = T.bn[0].hi;
        const uint32_t d = emu_collapse(T, tlas_cap + b, s->nodes.data(), &wide_count, &tri_count, tri_obj.data(), s->tris.data(), nullptr);
        blas_depth = std::max(blas_depth, d);
    }
    if (tri_count != n) return -1;
    s->nodes.resize(wide_count);
    const int rc = emu_rebuild_tlas(s, blo, bhi);
    s->depth = s->tlas_depth + blas_depth;
    return rc;
}

uint32_t emu_node_count(EmuScene *s) { return (uint32_t)s->nodes.size(); }
uint32_t emu_depth(EmuScene *s) { return s->depth; }
int emu_max_stack(EmuScene *s) { return s->max_stack; }
float emu_sah(EmuScene *s, int which) { return which ? s->sah_final : s->sah_lbvh; }
void emu_read_nodes(EmuScene *s, void *out) { memcpy(out, s->nodes.data(), s->nodes.size() * sizeof(Node8)); }
void emu_read_tris(EmuScene *s, void *out) { memcpy(out, s->tris.data(), s->tris.size() * sizeof(Tri48)); }

void emu_trace_rays(EmuScene *s, const float *rays, uint32_t n, uint32_t *hits, float *t_out, uint64_t *counters) {
    uint64_t nodes = 0, tris = 0;
    int max_stack = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, tris) reduction(max : max_stack)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Ray r;
        r.o = f3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]); r.tmin = rays[8 * i + 3];
        r.d = f3(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6]); r.tmax = rays[8 * i + 7];
        Hit h;
        EmuStack st;
        TraceCounters c = { 0, 0 };
        if (s->two_level) trace_closest_2l<true>((const uint4 *)s->nodes.data(), (const float4 *)s->tris.data(), (const float4 *)s->inst_leaves.data(), r, h, st, &c);
        else trace_closest<true>((const uint4 *)s->nodes.data(), (const float4 *)s->tris.data(), r, h, st, &c);
        hits[4 * i] = h.inst; hits[4 * i + 1] = h.prim; hits[4 * i + 2] = f2u(h.u); hits[4 * i + 3] = f2u(h.v);
        if (t_out) t_out[i] = h.inst != SOLB_MISS ? h.t : 0.0f;
        nodes += c.nodes; tris += c.tris;
        if (st.max_sp > max_stack) max_stack = st.max_sp;
    }
    if (counters) { counters[0] += nodes; counters[1] += tris; }
    if (max_stack > s->max_stack) s->max_stack = max_stack;
}

static void fill_fc(FrameConsts &fc, const float *uniforms, uint32_t w, uint32_t h) {
    memset(&fc, 0, sizeof(fc));
    memcpy(fc.view_inv, uniforms + 32, 64);
    memcpy(fc.proj_inv, uniforms + 64, 64);
    fc.origin = f3(fc.view_inv[12], fc.view_inv[13], fc.view_inv[14]);
    fc.tmin = fmaxf(1.0f, length(fc.origin)) * 1e-3f;
    fc.tmax = 10000.0f;
    fc.width = w; fc.height = h; fc.row_begin = 0; fc.band_rows = h; fc.band_stride = 0; fc.n_bands = 1;
    fc.frame = ((const uint32_t *)uniforms)[98];
}

void emu_debug(EmuScene *s, const float *uniforms, uint32_t w, uint32_t h, uint32_t *render, uint32_t *ids) {
    FrameConsts fc;
    fill_fc(fc, uniforms, w, h);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < (int64_t)h; y++)
        for (uint32_t x = 0; x < w; x++) {
            Ray r;
            r.o = fc.origin; r.d = primary_dir(fc, (float)x + 0.5f, (float)y + 0.5f); r.tmin = 0.001f; r.tmax = 1000.0f;
            Hit hit;
            EmuStack st;
            if (s->two_level) trace_closest_2l<false>((const uint4 *)s->nodes.data(), (const float4 *)s->tris.data(), (const float4 *)s->inst_leaves.data(), r, hit, st, (TraceCounters *)nullptr);
            else trace_closest<false>((const uint4 *)s->nodes.data(), (const float4 *)s->tris.data(), r, hit, st, (TraceCounters *)nullptr);
            float3 hv = r.d;
            if (hit.inst != SOLB_MISS) hv = f3(1.0f - hit.u - hit.v, hit.u, hit.v);
            const size_t p = (size_t)y * w + x;
            if (render) render[p] = pack_rgba8(hv.x, hv.y, hv.z, 0.0f);
            if (ids) { ids[2 * p] = hit.inst; ids[2 * p + 1] = hit.prim; }
        }
}

// mirrors solb_scene_set_textures: inst_tex[i] = texture of instance i or 0xffffffff
void emu_scene_set_textures(EmuScene *s, uint32_t n_tex, const uint8_t *const *rgba8, const uint32_t *width, const uint32_t *height,
                            const uint32_t *wrap_s, const uint32_t *wrap_t, const uint32_t *inst_tex) {
    float lut[256];
    for (int i = 0; i < 256; i++) {
        const float x = (float)i / 255.0f;
        lut[i] = x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f);
    }
    s->texels.assign(n_tex, {});
    s->tex.assign(n_tex, TexDesc{});
    for (uint32_t t = 0; t < n_tex; t++) {
        const size_t n = (size_t)width[t] * height[t];
        s->texels[t].resize(n);
        for (size_t i = 0; i < n; i++)
            s->texels[t][i] = make_float4(lut[rgba8[t][4 * i]], lut[rgba8[t][4 * i + 1]], lut[rgba8[t][4 * i + 2]], (float)rgba8[t][4 * i + 3] / 255.0f);
        s->tex[t].texels = s->texels[t].data();
        s->tex[t].width = width[t]; s->tex[t].height = height[t];
        s->tex[t].wrap_s = wrap_s[t] ? wrap_s[t] : 10497u; s->tex[t].wrap_t = wrap_t[t] ? wrap_t[t] : 10497u;
    }
    for (size_t i = 0; i < s->inst.size(); i++) {
        const uint32_t t1 = (n_tex && inst_tex[i] < n_tex) ? inst_tex[i] + 1u : 0u;
        memcpy(&s->inst[i].mat[10], &t1, sizeof(t1));
    }
}

// mirrors k_pathtrace_mega
void emu_pathtrace_frame(EmuScene *s, const float *uniforms, uint32_t w, uint32_t h, int accum_start, int enable_sky, int spp,
                         int max_bounces, int accum_mode, float *accum, uint32_t *render, uint64_t *stats) {
    FrameConsts fc;
    fill_fc(fc, uniforms, w, h);
    fc.accum_start = accum_start; fc.enable_sky = (uint32_t)enable_sky; fc.spp = (uint32_t)spp; fc.max_bounces = (uint32_t)max_bounces;
    fc.accum_mode = (uint32_t)accum_mode;
    fc.texb.tex = s->tex.data(); fc.texb.vertices = s->vertices.data(); fc.texb.indices = s->indices.data(); fc.texb.n_tex = (uint32_t)s->tex.size();
    uint64_t nrays = 0, nhits = 0, nnodes = 0, ntris = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : nrays, nhits, nnodes, ntris)
    for (int64_t y = 0; y < (int64_t)h; y++)
        for (uint32_t x = 0; x < w; x++) {
            uint32_t rng = tea(x + (uint32_t)y * w, fc.frame);
            float3 pixel = f3(0, 0, 0), thr = f3(1, 1, 1);
            uint32_t sample = 0, depth = 0;
            Ray r;
            r.tmin = fc.tmin; r.tmax = fc.tmax;
            { const float jx = next_rand(rng), jy = next_rand(rng); r.o = fc.origin; r.d = primary_dir(fc, (float)x + jx, (float)y + jy); }
            while (sample < fc.spp) {
                Hit hit;
                EmuStack st;
                TraceCounters ctr = { 0, 0 };
                if (s->two_level) trace_closest_2l<true>((const uint4 *)s->nodes.data(), (const float4 *)s->tris.data(), (const float4 *)s->inst_leaves.data(), r, hit, st, &ctr);
                else trace_closest<true>((const uint4 *)s->nodes.data(), (const float4 *)s->tris.data(), r, hit, st, &ctr);
                nnodes += ctr.nodes; ntris += ctr.tris;
                nrays++;
                bool end_path;
                if (hit.inst != SOLB_MISS) {
                    nhits++;
                    float3 hv;
                    const bool done = shade_hit(s->inst.data(), s->shade.data(), fc.texb, hit.inst, hit.gtri, hit.u, hit.v, r.o, r.d, rng, hv);
                    depth++;
                    thr = thr * hv;
                    end_path = done;
                    if (!done && depth > fc.max_bounces) { thr = f3(0, 0, 0); end_path = true; }
                } else { thr = thr * shade_miss(fc.enable_sky, r.d); end_path = true; }
                if (end_path) {
                    pixel = pixel + thr;
                    sample++;
                    if (sample < fc.spp) {
                        const float jx = next_rand(rng), jy = next_rand(rng);
                        r.o = fc.origin; r.d = primary_dir(fc, (float)x + jx, (float)y + jy);
                        thr = f3(1, 1, 1); depth = 0;
                    }
                }
            }
            const size_t p = (size_t)y * w + x;
            uint32_t rgba;
            const float4 old = make_float4(accum[4 * p], accum[4 * p + 1], accum[4 * p + 2], accum[4 * p + 3]);
            const float4 out = resolve_pixel(fc, pixel, old, rgba);
            accum[4 * p] = out.x; accum[4 * p + 1] = out.y; accum[4 * p + 2] = out.z; accum[4 * p + 3] = out.w;
            if (render) render[p] = rgba;
        }
    if (stats) { stats[0] += nrays; stats[1] += nhits; stats[2] += nnodes; stats[3] += ntris; }
}

// Node encoding self-test: random child sets through encode_node8, random slab results through assemble_hit_mask for all eight
// octants, against the mask built child by child from the ChildRefs.  Returns the number of mismatches.
uint32_t emu_mask_selftest(uint32_t seed, uint32_t rounds) {
    uint32_t bad = 0, rng = seed;
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    for (uint32_t r = 0; r < rounds; r++) {
        ChildRef ch[8];
        uint32_t n_tris = 0;
        for (int i = 0; i < 8; i++) {
            const uint32_t kind = rnd() % 4u;  // 0 empty, 1 internal, 2 / 3 leaf
            ch[i].valid = kind != 0u;
            ch[i].is_inner = kind == 1u;
            ch[i].lo = f3(0.1f * i, 0.0f, 0.0f); ch[i].hi = f3(0.1f * i + 0.05f, 1.0f, 1.0f);
            ch[i].tri_offset = 0; ch[i].tri_count = 0;
            if (kind >= 2u) { ch[i].tri_count = 1u + rnd() % (uint32_t)SOLB_MAX_LEAF_TRIS; ch[i].tri_offset = n_tris; n_tris += ch[i].tri_count; }
        }
        Node8 node;
        const uint32_t tri_base = rnd() & SOLB_TRI_BASE_MASK;
        encode_node8(node, f3(0, 0, 0), f3(1, 1, 1), rnd(), tri_base, ch);
        if ((node.q[1].y & SOLB_TRI_BASE_MASK) != tri_base) bad++;
        for (int i = 0; i < 8; i++) {  // decode_child_kind must return what was encoded
            uint32_t off, cnt;
            const int kind = decode_child_kind(node, i, off, cnt);
            const int want = !ch[i].valid ? 0 : (ch[i].is_inner ? 1 : 2);
            if (kind != want || (kind == 2 && (off != ch[i].tri_offset || cnt != ch[i].tri_count))) bad++;
        }
        for (uint32_t oct = 0; oct < 8; oct++) {
            TravRay tr;
            set_trav_octant(tr, f3(oct & 4u ? -1.0f : 1.0f, oct & 2u ? -1.0f : 1.0f, oct & 1u ? -1.0f : 1.0f));
            const uint32_t hits = rnd() & 0xffu;
            uint32_t hb[2] = { 0, 0 }, want = 0;
            for (int i = 0; i < 8; i++) {
                if (!((hits >> i) & 1u)) continue;
                hb[i >> 2] |= 1u << (8 * (i & 3));
                if (!ch[i].valid) continue;
                if (ch[i].is_inner) want |= 1u << (24 + ((uint32_t)i ^ tr.oct_inv));
                else want |= ((1u << ch[i].tri_count) - 1u) << ch[i].tri_offset;
            }
            if (assemble_hit_mask(hb[0], hb[1], node.q[0].w >> 24, node.q[1], tr.pow4_lo, tr.pow4_hi) != want) bad++;
        }
    }
    return bad;
}

uint32_t emu_tea(uint32_t a, uint32_t b) { return tea(a, b); }
float emu_next_rand(uint32_t *rng) { return next_rand(*rng); }

}  // extern "C"
