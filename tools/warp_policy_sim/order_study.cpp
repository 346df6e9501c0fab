// Exploration: how many node visits / triangle tests would exact front-to-back child ordering save over octant order?
#include <algorithm>
#include <cstdio>
#include <vector>
#include "build.cuh"
using namespace solb;
struct Ent { uint32_t node; float t; };
static float g_key[8]; static int g_mode = 0;
// decode child boxes and intersect in f32 (same conservative boxes as intersect_node)
static int children_hit(const Node8 &n, float3 o, float3 idir, float tmin, float tmax, uint32_t *child_node, float *child_t, uint32_t *leaf_base, uint32_t *leaf_cnt, float *leaf_t, int &n_leaf) {
    const uint32_t *w = (const uint32_t *)&n;
    const uint32_t imask = w[3] >> 24, child_base = w[4], tri_base = w[5] & SOLB_TRI_BASE_MASK;
    int ni = 0; n_leaf = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t l_off, l_cnt;
        const int kind = decode_child_kind(n, i, l_off, l_cnt);
        if (!kind) continue;
        float3 lo, hi; decode_child_box(n, i, lo, hi);
        float t0 = tmin, t1 = tmax;
        const float lo3[3] = { lo.x, lo.y, lo.z }, hi3[3] = { hi.x, hi.y, hi.z }, o3[3] = { o.x, o.y, o.z }, id3[3] = { idir.x, idir.y, idir.z };
        for (int a = 0; a < 3; a++) {
            float ta = (lo3[a] - o3[a]) * id3[a], tb = (hi3[a] - o3[a]) * id3[a];
            if (ta > tb) std::swap(ta, tb);
            ta -= fabsf(ta) * 1e-6f + 1e-7f; tb += fabsf(tb) * 1e-6f + 1e-7f;
            t0 = fmaxf(t0, ta); t1 = fminf(t1, tb);
        }
        if (t0 > t1) continue;
        if (kind == 1) { child_node[ni] = child_base + popc32(imask & ((1u << i) - 1)); child_t[ni] = t0;
            const float sx = idir.x < 0 ? -1.f : 1.f, sy = idir.y < 0 ? -1.f : 1.f, sz = idir.z < 0 ? -1.f : 1.f;
            g_key[ni] = g_mode == 7 ? sx * (lo.x + hi.x) + sy * (lo.y + hi.y) + sz * (lo.z + hi.z)            // centre on the octant diagonal
                      : g_mode == 8 ? sx * (idir.x < 0 ? hi.x : lo.x) + sy * (idir.y < 0 ? hi.y : lo.y) + sz * (idir.z < 0 ? hi.z : lo.z)  // near corner on the diagonal
                      : 0.f;
            ni++; }
        else { leaf_base[n_leaf] = tri_base + l_off; leaf_cnt[n_leaf] = l_cnt; leaf_t[n_leaf] = t0; n_leaf++; }
    }
    return ni;
}
extern "C" void order_sim(const Node8 *nodes, const float4 *tris, const float *rays, uint32_t n, int sorted, double *out) {
    double nn = 0, nt = 0; g_mode = sorted;
    for (uint32_t r = 0; r < n; r++) {
        const float *p = rays + 8 * (size_t)r;
        const float3 o = f3(p[0], p[1], p[2]), d = f3(p[4], p[5], p[6]);
        const float tmin = p[3]; float tmax = p[7];
        const float3 idir = f3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
        const RayFrame fr = make_ray_frame(d);
        std::vector<Ent> st; st.push_back({ 0, tmin });
        while (!st.empty()) {
            Ent e = st.back(); st.pop_back();
            if (e.t > tmax) continue;   // culled by a closer hit found meanwhile (only helps with distance info)
            nn++;
            uint32_t cn[8], lb[8], lc[8]; float ct[8], lt[8]; int nl;
            int ni = children_hit(nodes[e.node], o, idir, tmin, tmax, cn, ct, lb, lc, lt, nl);
            // leaves first (as the kernel does)
            int lo_[8]; for (int i = 0; i < nl; i++) lo_[i] = i;
            if (sorted == 1) std::sort(lo_, lo_ + nl, [&](int a, int b) { return lt[a] < lt[b]; });
            for (int k = 0; k < nl; k++) { int i = lo_[k]; if (sorted >= 1 && sorted <= 2 && lt[i] > tmax) continue; for (uint32_t j = 0; j < lc[i]; j++) {
                nt++; const float4 *tp = tris + (size_t)(lb[i] + j) * 3; float t, u, v;
                if (intersect_tri(o, d, fr, xyz(tp[0]), xyz(tp[1]), xyz(tp[2]), tmin, tmax, t, u, v)) tmax = t; } }
            int io[8]; for (int i = 0; i < ni; i++) io[i] = i;
            if (sorted == 1) std::sort(io, io + ni, [&](int a, int b) { return ct[a] > ct[b]; });   // push far first
            if (sorted == 7 || sorted == 8) std::sort(io, io + ni, [&](int a, int b) { return g_key[a] > g_key[b]; });  // push far first
            if (sorted == 5 || sorted == 6) {  // nearest child visited first, the rest keep their slot order
                int best = -1; for (int i = 0; i < ni; i++) if (best < 0 || ct[i] < ct[best]) best = i;
                if (best >= 0) { int k2 = 0; int tmp[8]; for (int i = 0; i < ni; i++) if (i != best) tmp[k2++] = i; tmp[k2++] = best; for (int i = 0; i < ni; i++) io[i] = tmp[i]; }
                // leaves: nearest leaf first too
                if (sorted == 6) { int bl = -1; for (int i = 0; i < nl; i++) if (bl < 0 || lt[i] < lt[bl]) bl = i; (void)bl; }
            }
            float gmin = 3.4e38f; for (int i = 0; i < ni; i++) gmin = fminf(gmin, ct[i]);
            for (int k = 0; k < ni; k++) {
                float tt = tmin;
                if (sorted == 1 || sorted == 2) tt = ct[io[k]];          // per-child distance
                else if (sorted == 3) tt = gmin;                          // one distance per pushed group (never refreshed)
                else if (sorted == 4) { tt = 3.4e38f; for (int q = 0; q <= k; q++) tt = fminf(tt, ct[io[q]]); }  // group min refreshed as children are consumed
                st.push_back({ cn[io[k]], tt });
            }
        }
    }
    out[0] = nn / n; out[1] = nt / n;
}
