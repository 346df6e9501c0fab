// Warp-scheduling simulator: replays real path-tracing rays (tunnel.gltf, grouped by wave) through lane-level traversal
// state machines under different warp policies and counts instruction slots.  Exploration only.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include "build.cuh"
#include "shade.cuh"
using namespace solb;

struct SimStack { uint2 e[128]; int sp = 0; void push(uint2 v) { e[sp++] = v; } uint2 pop() { return e[--sp]; } bool empty() const { return sp == 0; } };
struct Lane {
    bool has_ray = false; TravRay tr; float tmax; Hit hit; uint2 ngroup, tgroup; SimStack st;
};
struct Costs { double node = 230, tri = 110, coop = getenv("SIM_COOP") ? atof(getenv("SIM_COOP")) : 180, overhead = 30, fetch = 160, pop = 14; };

extern "C" void warp_sim(const uint4 *nodes, const float4 *tris, const float *rays, uint32_t n, int policy, int fetch_idle, int block_thresh,
                         double *out /* [0] total slots, [1] node steps, [2] node lanes, [3] tri steps, [4] tri lanes, [5] coop rounds, [6] coop tests, [7] iterations */) {
    Costs C;
    const uint32_t rays_per_warp = 438;
    const uint32_t n_warps = std::max<uint32_t>(1, n / rays_per_warp);
    std::vector<std::vector<Lane>> W(n_warps, std::vector<Lane>(32));
    std::vector<uint32_t> pool_next(n_warps, 0), pool_end(n_warps, 0);
    std::vector<char> exhausted(n_warps, 0), done(n_warps, 0);
    uint32_t head = 0, n_done = 0;
    double slots = 0, node_steps = 0, node_lanes = 0, tri_steps = 0, tri_lanes = 0, coop_rounds = 0, coop_tests = 0, iters = 0;
    while (n_done < n_warps) {
        for (uint32_t w = 0; w < n_warps; w++) {
            if (done[w]) continue;
            auto &L = W[w];
            int idle = 0; for (auto &l : L) idle += !l.has_ray;
            if (!exhausted[w] && idle >= fetch_idle) {
                for (auto &l : L) {
                    if (l.has_ray) continue;
                    if (pool_next[w] >= pool_end[w]) {
                        if (head >= n) { exhausted[w] = 1; break; }
                        pool_next[w] = head; pool_end[w] = std::min(head + 128, n); head += 128;
                    }
                    const float *r = rays + 8 * (size_t)pool_next[w]++;
                    l.tr = make_trav_ray(f3(r[0], r[1], r[2]), f3(r[4], r[5], r[6]), r[3]);
                    l.tmax = r[7]; l.hit.inst = SOLB_MISS; l.ngroup = SOLB_ROOT_GROUP; l.tgroup = make_uint2(0, 0); l.st.sp = 0; l.has_ray = true;
                }
                slots += C.fetch;
            }
            int active = 0; for (auto &l : L) active += l.has_ray;
            if (!active) { done[w] = 1; n_done++; continue; }
            iters++;
            int nn = 0, nt = 0, blocked = 0, tests = 0;
            for (auto &l : L) if (l.has_ray) {
                const bool wn = l.ngroup.y & 0xff000000u, wt = l.tgroup.y != 0;
                nn += wn; nt += wt; blocked += (wt && !wn); tests += popc32(l.tgroup.y);
            }
            slots += C.overhead;
            if (policy == 0) {
                if (nt > 0 && nt >= nn) {
                    for (auto &l : L) if (l.has_ray && l.tgroup.y) trav_tri_step(tris, l.tr, l.tmax, l.tgroup, l.hit);
                    slots += C.tri; tri_steps++; tri_lanes += nt;
                } else if (nn > 0) {
                    for (auto &l : L) if (l.has_ray && (l.ngroup.y & 0xff000000u)) { if (l.tgroup.y) l.st.push(l.tgroup); trav_node_step(nodes, l.tr, l.tmax, l.ngroup, l.tgroup, l.st); }
                    slots += C.node; node_steps++; node_lanes += nn;
                }
            } else if (policy == 2) {
                // on trigger: plain tri steps repeated until no lane holds triangle work
                if (nt > 0 && (blocked >= block_thresh || nn == 0)) {
                    for (;;) {
                        int k = 0;
                        for (auto &l : L) if (l.has_ray && l.tgroup.y) { trav_tri_step(tris, l.tr, l.tmax, l.tgroup, l.hit); k++; }
                        if (!k) break;
                        slots += C.tri + 6; tri_steps++; tri_lanes += k;
                    }
                } else if (nn > 0) {
                    for (auto &l : L) if (l.has_ray && (l.ngroup.y & 0xff000000u)) { if (l.tgroup.y) l.st.push(l.tgroup); trav_node_step(nodes, l.tr, l.tmax, l.ngroup, l.tgroup, l.st); }
                    slots += C.node; node_steps++; node_lanes += nn;
                }
            } else {
                const int tests_thresh = getenv("SIM_TESTS") ? atoi(getenv("SIM_TESTS")) : 32;
                const int nn_min = getenv("SIM_NNMIN") ? atoi(getenv("SIM_NNMIN")) : 0;
                if (nt > 0 && (blocked >= block_thresh || nn == 0 || tests >= tests_thresh || nn < nn_min)) {
                    for (auto &l : L) if (l.has_ray) while (l.tgroup.y) trav_tri_step(tris, l.tr, l.tmax, l.tgroup, l.hit);
                    const int rounds = (tests + 31) / 32;
                    slots += C.coop * rounds; coop_rounds += rounds; coop_tests += tests;
                } else if (nn > 0) {
                    for (auto &l : L) if (l.has_ray && (l.ngroup.y & 0xff000000u)) { if (l.tgroup.y) l.st.push(l.tgroup); trav_node_step(nodes, l.tr, l.tmax, l.ngroup, l.tgroup, l.st); }
                    slots += C.node; node_steps++; node_lanes += nn;
                }
            }
            for (auto &l : L) if (l.has_ray && !(l.ngroup.y & 0xff000000u) && !l.tgroup.y) {
                if (l.st.empty()) l.has_ray = false;
                else { const uint2 e = l.st.pop(); if (e.y & 0xff000000u) l.ngroup = e; else l.tgroup = e; }
            }
            slots += C.pop;
        }
    }
    out[0] = slots; out[1] = node_steps; out[2] = node_lanes; out[3] = tri_steps; out[4] = tri_lanes; out[5] = coop_rounds; out[6] = coop_tests; out[7] = iters;
}
