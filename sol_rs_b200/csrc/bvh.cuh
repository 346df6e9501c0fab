// bvh.cuh — 8-wide compressed BVH: node/triangle layouts, watertight ray/triangle test, traversal.
//
// Replaces what the reference delegates to the Vulkan driver: traceRayEXT
// (assets/glsl/pathtrace.rgen:65-76, ao.rgen:59-70, debug.rgen:34) over the TLAS/BLAS built in
// src/ray/acceleration.rs.  Contract restated: closest hit, opaque, two-sided (cull disabled,
// acceleration.rs:337), mask 0xFF, tmin < t < tmax, attribs = barycentrics of v1, v2.
//
// Layout (after Ylitie, Karras, Laine 2017, "Efficient Incoherent Ray Traversal on GPUs Through
// Compressed Wide BVHs"): 80-byte nodes = 5 x 16-byte vector loads; 48-byte triangles = 3 x float4.
#pragma once
#include "common.cuh"

namespace solb {

// ---- 80-byte node, addressed as 5 uint4 --------------------------------------------------------
//  q0: px, py, pz (f32 bits), [ex | ey<<8 | ez<<16 | imask<<24]
//  q1: child_base, tri_base | count0<<28, tmask[0..3], tmask[4..7]
//  q2: qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7]
//  q3: qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7]
//  q4: qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7]
// imask bit i: slot i holds an internal child (children are numbered from child_base in slot order).
// tmask[i]: 0 for internal and empty slots; for a leaf of 1 or 2 triangles its static contribution to the triangle hit mask,
//   (unary count = 1 or 3) << (offset inside its group of four slots).  Triangle offsets are cumulative in slot order, so the
//   leaves of slots 0-3 own triangle bits [0, count0) and those of slots 4-7 the bits from count0 up: a group holds at most
//   4 x 2 triangles and every contribution fits one byte.  With the eight slab-test results packed as bytes (1 = hit) the
//   triangle hit mask is two dot products (IDP.4A, FMA-heavy pipe) instead of two field extractions, a variable shift and an OR per
//   child on the ALU pipe (Ylitie et al.'s meta bytes, used until session 3: 56 of the node step's 237 instructions).
// child box i = p + q * 2^(e-127) per axis, conservative (floor / ceil).
struct Node8 {
    uint4 q[5];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");
#define SOLB_TRI_BASE_MASK 0x0fffffffu  // q1.y: triangle base (28 bits) | count0 << 28

// ---- 48-byte triangle: world-space vertices; w lanes carry the ids the hit shaders need ----------
//  v0.w = gl_InstanceID, v1.w = gl_PrimitiveID, v2.w = global triangle ordinal (index of the shading record)
struct Tri48 {
    float4 v0, v1, v2;
};
static_assert(sizeof(Tri48) == 48, "Tri48 must be 48 bytes");

#define SOLB_MAX_LEAF_TRIS 2  // a leaf's mask contribution must fit one byte (above); 3 buys nothing on the shipped scenes
#ifndef SOLB_SM_STACK
#define SOLB_SM_STACK 8      // per-lane entries kept in shared memory
#endif
#define SOLB_LOCAL_STACK (64 - SOLB_SM_STACK)  // spill (local memory)
// The warp-cooperative kernel may park two entries per level (the siblings' node group and a postponed triangle
// group), a two-level walk adds one sentinel: 2 * (TLAS depth + BLAS depth) + 1 <= SOLB_SM_STACK + SOLB_LOCAL_STACK.
#define SOLB_MAX_WIDE_DEPTH ((SOLB_SM_STACK + SOLB_LOCAL_STACK - 2) / 2)

struct Ray {
    float3 o;
    float tmin;
    float3 d;
    float tmax;
};

struct Hit {
    float t, u, v;     // u, v = hitAttributeEXT attribs.xy (weights of vertex 1 and 2)
    uint32_t inst;     // gl_InstanceID, SOLB_MISS on miss
    uint32_t prim;     // gl_PrimitiveID
    uint32_t gtri;     // global triangle ordinal
};

struct TraceCounters {
    uint32_t nodes, tris;
};

// ---- watertight ray/triangle test ------------------------------------------------------------------
// Woop, Benthin, Wald 2013 ("Watertight Ray/Triangle Intersection") with the axis-permuting shear
// replaced by a per-ray orthonormal frame (e1, e2, d): every vertex is projected to 2D ray space with
// the same explicit fma sequence, so a vertex shared by two triangles lands on bit-identical
// coordinates, and the 2D edge functions are evaluated without FMA contraction so the value of a shared
// edge is exactly negated in the neighbour: a ray cannot slip between triangles that share vertices.
// Exact zeros are re-evaluated in f64.  Barycentric error ~ eps * distance / triangle size.
struct RayFrame {
    float3 e1, e2;
};

SOLB_HD RayFrame make_ray_frame(float3 d) {
    // Duff et al. 2017 branchless orthonormal basis around normalize(d)
    const float inv = fast_rsqrt(dot(d, d));
    const float3 n = f3(d.x * inv, d.y * inv, d.z * inv);
    const float s = (n.z < 0.0f ? -1.0f : 1.0f);
    const float a = -fast_rcp(s + n.z);
    const float b = n.x * n.y * a;
    RayFrame f;
    f.e1 = f3(1.0f + s * n.x * n.x * a, s * b, -s * n.x);
    f.e2 = f3(b, s + n.y * n.y * a, -n.y);
    return f;
}

SOLB_HD float proj_rn(float3 a, float3 e) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a.z, e.z, __fmaf_rn(a.y, e.y, __fmul_rn(a.x, e.x)));
#else
    return fmaf(a.z, e.z, fmaf(a.y, e.y, mul_rn(a.x, e.x)));
#endif
}
SOLB_HD float edge2d(float ax, float ay, float bx, float by) { return sub_rn(mul_rn(ax, by), mul_rn(ay, bx)); }

SOLB_HD bool intersect_tri(float3 o, float3 d, const RayFrame &fr, float3 p0, float3 p1, float3 p2, float tmin, float tmax,
                           float &t_out, float &u_out, float &v_out) {
    const float3 A = f3(sub_rn(p0.x, o.x), sub_rn(p0.y, o.y), sub_rn(p0.z, o.z));
    const float3 B = f3(sub_rn(p1.x, o.x), sub_rn(p1.y, o.y), sub_rn(p1.z, o.z));
    const float3 C = f3(sub_rn(p2.x, o.x), sub_rn(p2.y, o.y), sub_rn(p2.z, o.z));
    const float ax = proj_rn(A, fr.e1), ay = proj_rn(A, fr.e2);
    const float bx = proj_rn(B, fr.e1), by = proj_rn(B, fr.e2);
    const float cx = proj_rn(C, fr.e1), cy = proj_rn(C, fr.e2);
    float U = edge2d(cx, cy, bx, by);  // weight of vertex 0
    float V = edge2d(ax, ay, cx, cy);  // weight of vertex 1
    float W = edge2d(bx, by, ax, ay);  // weight of vertex 2
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        const double Ud = (double)cx * by - (double)cy * bx;
        const double Vd = (double)ax * cy - (double)ay * cx;
        const double Wd = (double)bx * ay - (double)by * ax;
        if ((Ud < 0.0 || Vd < 0.0 || Wd < 0.0) && (Ud > 0.0 || Vd > 0.0 || Wd > 0.0)) return false;
        U = (float)Ud; V = (float)Vd; W = (float)Wd;
    } else if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) {
        return false;  // two-sided: all three must share a sign
    }
    const float det = U + V + W;
    if (det == 0.0f) return false;  // edge-on triangle
    const float rdet = fast_rcp(det);
    // Depth from the plane through the three vertices (n = e1 x e2 from exact-ish edge differences):
    // t = (p0 - o).n / d.n.  On the long sliver triangles of tunnel.gltf (9.5 x 0.05 units) this is ~5x
    // more accurate than interpolating vertex depths with the barycentrics, and as accurate as the f32
    // rounding of the stored world-space vertices allows (~1e-4 there).
    const float3 e1 = p1 - p0, e2 = p2 - p0;
    const float3 n = cross(e1, e2);
    const float den = dot(d, n);
    if (den == 0.0f) return false;
    const float t = fast_div(dot(A, n), den);
    if (!(t > tmin && t <= tmax)) return false;  // t == tmax: a tie with the hit in hand, settled by the caller (trav_tri_step)
    t_out = t;
    u_out = V * rdet;
    v_out = W * rdet;
    return true;
}

// ---- node test: returns hit mask (bits 31..24 internal children in traversal priority order,
//      bits 23..0 triangles of the hit leaf children) -------------------------------------------------
// Dequantisation without I2F (quarter-rate XU pipe; it was the top pipe of the first traversal kernel,
// profiles/r01_ncu_k_wf_trace_a.txt) and without a separate subtraction: byte j of w is dropped into the mantissa of the
// float 32768.0, giving m = 32768 + s q exactly; the 32768 is folded into the per-node offset, t = m * a + (b - 32768 a).
// Folding costs at most |a| / 512 of rounding (a = one quantisation cell in units of t), which the near / far offsets below
// absorb conservatively.  Two encodings of the same idea:
//   PRMT (SOLB_Q2M_PRMT): one byte permute puts q into mantissa bits [15:8]: m = 32768 + q, s = 1.  PRMT runs on the ALU pipe,
//     which issues one warp-instruction every two cycles; with 48 of them on top of the FMNMX / LOP3 / ISETP of the slab test
//     the node step was bound by that pipe (~120 ALU instructions = 240 cycles against ~185 issue slots).
//   IDP.4A (default): dp4a(w, 0xff << 8j, bits(32768.0f)) = bits + 255 q: m = 32768 + q * 255/256, s = 255/256, and the
//     per-node scale a carries the factor 256/255.  Same instruction count, but IDP.4A issues on the FMA-heavy pipe
//     (tools/microbench/pipes.cu on B200: PRMT 0.48, IDP.4A 0.48, FFMA 0.89 per clock alone; PRMT + IDP.4A interleaved 0.93 per
//     clock, i.e. different pipes), so the ALU pipe sheds 40 % of its node-step load.
#if defined(__CUDACC__)
// bit pattern of 32768.0f in (non-const) constant memory: opaque to the optimiser, so PRMT / IDP take it as the register /
// constant operand and the byte selector as an immediate
__constant__ uint32_t c_q2m_bias = 0x47000000u;
#endif
#if defined(SOLB_Q2M_PRMT)
#define SOLB_Q2M_SCALE 1.0f
#else
#define SOLB_Q2M_SCALE 1.00392163f  // 256 / 255 rounded up: m * (a * 256/255) = (32768 * 256/255 + q) * a
#endif
SOLB_HD float q2m(uint32_t w, int j) {
#if defined(__CUDA_ARCH__)
    // selector as an immediate, the 32768.0f pattern as the register operand: otherwise ptxas keeps re-materialising
    // the four selectors in registers (one extra IMAD.U32 per PRMT in the first build)
    uint32_t r;
    const uint32_t k = c_q2m_bias;
#if defined(SOLB_Q2M_PRMT)
    switch (j) {
        case 0: asm("prmt.b32 %0, %1, %2, 0x7604;" : "=r"(r) : "r"(w), "r"(k)); break;
        case 1: asm("prmt.b32 %0, %1, %2, 0x7614;" : "=r"(r) : "r"(w), "r"(k)); break;
        case 2: asm("prmt.b32 %0, %1, %2, 0x7624;" : "=r"(r) : "r"(w), "r"(k)); break;
        default: asm("prmt.b32 %0, %1, %2, 0x7634;" : "=r"(r) : "r"(w), "r"(k)); break;
    }
#else
    switch (j) {
        case 0: asm("dp4a.u32.u32 %0, %1, 0x000000ff, %2;" : "=r"(r) : "r"(w), "r"(k)); break;
        case 1: asm("dp4a.u32.u32 %0, %1, 0x0000ff00, %2;" : "=r"(r) : "r"(w), "r"(k)); break;
        case 2: asm("dp4a.u32.u32 %0, %1, 0x00ff0000, %2;" : "=r"(r) : "r"(w), "r"(k)); break;
        default: asm("dp4a.u32.u32 %0, %1, 0xff000000, %2;" : "=r"(r) : "r"(w), "r"(k)); break;
    }
#endif
    return __uint_as_float(r);
#elif defined(SOLB_Q2M_PRMT)
    return u2f(0x47000000u | (((w >> (8 * j)) & 0xffu) << 8));
#else
    return u2f(0x47000000u + 255u * ((w >> (8 * j)) & 0xffu));
#endif
}

// a . b over the four bytes + c (IDP.4A.U8.U8)
SOLB_HD uint32_t dot4_u8(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __dp4a(a, b, c);
#else
    for (int k = 0; k < 4; k++) c += ((a >> (8 * k)) & 0xffu) * ((b >> (8 * k)) & 0xffu);
    return c;
#endif
}

// Hit mask of a node from the byte-packed slab results (hb_lo / hb_hi: byte j = 1 if slot j / 4 + j passed).
//   bits 31..24: internal children, slot s -> bit 24 + (s ^ oct_inv), i.e. front-to-back along the ray's octant when popped high
//                bit first; pow4_lo / pow4_hi = the ray's 1 << (s ^ oct_inv) for s = 0..3 / 4..7, one byte each
//   bits 23..0 : triangles of the hit leaves = their static contributions (tmask bytes), slots 4-7 shifted past the bits of slots 0-3
SOLB_HD uint32_t assemble_hit_mask(uint32_t hb_lo, uint32_t hb_hi, uint32_t imask, const uint4 q1, uint32_t pow4_lo, uint32_t pow4_hi) {
    const uint32_t inner_lo = ((imask & 0xfu) * 0x00204081u) & 0x01010101u;  // imask bits 0-3 spread to bytes
    const uint32_t inner_hi = ((imask >> 4) * 0x00204081u) & 0x01010101u;
    const uint32_t inner_hits = dot4_u8(hb_lo & inner_lo, pow4_lo, dot4_u8(hb_hi & inner_hi, pow4_hi, 0u));
    const uint32_t t_lo = dot4_u8(hb_lo, q1.z, 0u), t_hi = dot4_u8(hb_hi, q1.w, 0u);
    return (inner_hits << 24) | t_lo | (t_hi << (q1.y >> 28));
}

SOLB_HD uint32_t intersect_node(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3, const uint4 q4,
                                float3 o, float3 idir, uint32_t pow4_lo, uint32_t pow4_hi, float tmin, float tmax) {
    const uint32_t e = q0.w;
    const float ax = u2f((e & 0xffu) << 23) * idir.x * SOLB_Q2M_SCALE;
    const float ay = u2f(((e >> 8) & 0xffu) << 23) * idir.y * SOLB_Q2M_SCALE;
    const float az = u2f(((e >> 16) & 0xffu) << 23) * idir.z * SOLB_Q2M_SCALE;
    const float bx = (u2f(q0.x) - o.x) * idir.x;
    const float by = (u2f(q0.y) - o.y) * idir.y;
    const float bz = (u2f(q0.z) - o.z) * idir.z;
    // offsets with the 32768 bias removed; near planes pulled in, far planes pushed out by |a| / 256
    const float cx = fmaf(-32768.0f, ax, bx), cy = fmaf(-32768.0f, ay, by), cz = fmaf(-32768.0f, az, bz);
    const float px = fabsf(ax) * 0.00390625f, py = fabsf(ay) * 0.00390625f, pz = fabsf(az) * 0.00390625f;
    const float nbx = cx - px, nby = cy - py, nbz = cz - pz;
    const float fbx = cx + px, fby = cy + py, fbz = cz + pz;
    uint32_t hb[2] = { 0u, 0u };  // slab-test results of slots 0-3 / 4-7, one byte each (1 = hit)
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const uint32_t lox = g ? q2.y : q2.x, loy = g ? q2.w : q2.z, loz = g ? q3.y : q3.x;
        const uint32_t hix = g ? q3.w : q3.z, hiy = g ? q4.y : q4.x, hiz = g ? q4.w : q4.z;
        // near/far planes per axis depend only on the ray's direction signs
        const uint32_t nx = idir.x < 0.0f ? hix : lox, fx = idir.x < 0.0f ? lox : hix;
        const uint32_t ny = idir.y < 0.0f ? hiy : loy, fy = idir.y < 0.0f ? loy : hiy;
        const uint32_t nz = idir.z < 0.0f ? hiz : loz, fz = idir.z < 0.0f ? loz : hiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float t0x = fmaf(q2m(nx, j), ax, nbx), t1x = fmaf(q2m(fx, j), ax, fbx);
            const float t0y = fmaf(q2m(ny, j), ay, nby), t1y = fmaf(q2m(fy, j), ay, fby);
            const float t0z = fmaf(q2m(nz, j), az, nbz), t1z = fmaf(q2m(fz, j), az, fbz);
            const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
            const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
            // (an empty slot decodes to the point box at the node origin; if a ray passes it, its tmask byte and imask bit are 0)
            if (cmin <= cmax) hb[g] |= 1u << (8 * j);
        }
    }
    return assemble_hit_mask(hb[0], hb[1], e >> 24, q1, pow4_lo, pow4_hi);
}

SOLB_HD float safe_rcp_dir(float d) {
    const float tiny = 1e-20f;
    return fast_rcp(fabsf(d) > tiny ? d : (f2u(d) >> 31 ? -tiny : tiny));
}

#if defined(__CUDA_ARCH__)
#define SOLB_LDG4(p) __ldg(p)
#else
#define SOLB_LDG4(p) (*(p))
#endif

// Per-ray constants of a traversal
struct TravRay {
    float3 o, d, idir;
    uint32_t oct_inv;           // 7 - octant: slot ^ oct_inv = traversal priority of a slot
    uint32_t pow4_lo, pow4_hi;  // byte j = 1 << ((4 g + j) ^ oct_inv) for g = 0 / 1
    RayFrame frame;
    float tmin;
};

// octant-dependent constants of a direction
SOLB_HD void set_trav_octant(TravRay &t, float3 d) {
    const uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
    t.oct_inv = 7u - oct;
    // 1 << (j ^ k) for j = 0..3 and the low two bits k of oct_inv, in nibble 0 or 1 of every byte by bit 2 of oct_inv
    const uint32_t k = t.oct_inv & 3u;
    const uint32_t p4 = k == 0u ? 0x08040201u : (k == 1u ? 0x04080102u : (k == 2u ? 0x02010804u : 0x01020408u));
    t.pow4_lo = (t.oct_inv & 4u) ? p4 << 4 : p4;
    t.pow4_hi = (t.oct_inv & 4u) ? p4 : p4 << 4;
}

SOLB_HD TravRay make_trav_ray(float3 o, float3 d, float tmin) {
    TravRay t;
    t.o = o; t.d = d; t.tmin = tmin;
    t.idir = f3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    set_trav_octant(t, d);
    t.frame = make_ray_frame(d);
    return t;
}

#define SOLB_ROOT_GROUP make_uint2(0u, 0x80000000u)  // "child 7^oct_inv of a virtual parent", base 0

// One node step.  Precondition: ngroup carries at least one internal-child hit.  Pops the nearest child,
// pushes the remaining siblings, intersects the child's 8 boxes and returns its node group in ngroup and
// its triangle group in tgroup.
template <class Stack>
SOLB_HD void trav_node_step(const uint4 *__restrict__ nodes, const TravRay &tr, float tmax, uint2 &ngroup, uint2 &tgroup, Stack &stack) {
    const uint32_t hits_imask = ngroup.y;
    const int child_bit = bfind32(hits_imask);
    const uint32_t child_base = ngroup.x;
    ngroup.y &= ~(1u << child_bit);
    if (ngroup.y & 0xff000000u) stack.push(ngroup);
    const uint32_t slot = (uint32_t)(child_bit - 24) ^ tr.oct_inv;
    const uint32_t rel = (uint32_t)popc32(hits_imask & ~(0xffffffffu << slot));
    const uint4 *np = nodes + (size_t)(child_base + rel) * 5;
    const uint4 q0 = SOLB_LDG4(np + 0), q1 = SOLB_LDG4(np + 1), q2 = SOLB_LDG4(np + 2), q3 = SOLB_LDG4(np + 3), q4 = SOLB_LDG4(np + 4);
    const uint32_t hm = intersect_node(q0, q1, q2, q3, q4, tr.o, tr.idir, tr.pow4_lo, tr.pow4_hi, tr.tmin, tmax);
    ngroup = make_uint2(q1.x, (hm & 0xff000000u) | (q0.w >> 24));
    tgroup = make_uint2(q1.y & SOLB_TRI_BASE_MASK, hm & 0x00ffffffu);
}

// One triangle step.  Precondition: tgroup.y != 0.  Tests the highest pending triangle of the group.
// Returns true when the hit record was replaced.  hit.gtri must hold the triangle of the hit in hand (SOLB_MISS: none).
// Closest hit is tmin < t < tmax; two triangles at exactly the same t (coplanar neighbours hit on their shared edge) are told
// apart by their global ordinal, so the result is a function of the ray alone: the ORDER in which a ray's candidates are
// tested depends on the other rays of its warp (the vote), and with "first tested wins" two runs of a 1080p frame differed
// in one path out of 50 million.
SOLB_HD bool trav_tri_step(const float4 *__restrict__ tris, const TravRay &tr, float &tmax, uint2 &tgroup, Hit &hit) {
    const int ti = bfind32(tgroup.y);
    tgroup.y &= ~(1u << ti);
    const float4 *tp = tris + (size_t)(tgroup.x + (uint32_t)ti) * 3;
    const float4 v0 = SOLB_LDG4(tp + 0), v1 = SOLB_LDG4(tp + 1), v2 = SOLB_LDG4(tp + 2);
    float t, u, v;
    if (intersect_tri(tr.o, tr.d, tr.frame, xyz(v0), xyz(v1), xyz(v2), tr.tmin, tmax, t, u, v) &&
        (t < tmax || (hit.gtri != SOLB_MISS && f2u(v2.w) < hit.gtri))) {
        tmax = t;
        hit.t = t; hit.u = u; hit.v = v;
        hit.inst = f2u(v0.w); hit.prim = f2u(v1.w); hit.gtri = f2u(v2.w);
        return true;
    }
    return false;
}

// ---- two-level traversal (TLAS of instances over shared, object-space BLASes) -------------------------
// The reference's TLAS/BLAS split (src/ray/acceleration.rs:344-400): a TLAS leaf holds 64-byte instance
// records instead of triangles.  Entering an instance transforms the ray into object space WITHOUT
// renormalising the direction, so t keeps its world-space meaning (tmin / tmax / closest hit carry over),
// saves the TLAS groups still in hand under a sentinel entry on the traversal stack and restarts at the
// BLAS root; popping the sentinel returns to world space.
//   r0, r1, r2 = rows of the world->object 3x4 matrix, r3 = (bits(BLAS root node), bits(instance id), 0, 0)
struct InstLeaf {
    float4 r0, r1, r2, r3;
};
static_assert(sizeof(InstLeaf) == 64, "InstLeaf must be 64 bytes");
#define SOLB_STACK_SENTINEL make_uint2(0xffffffffu, 0u)  // .y == 0 never occurs in a pushed node / triangle group

// One instance step (TLAS mode).  Precondition: tgroup.y != 0.  world_o / world_d = the ray as traced.
template <class Stack>
SOLB_HD void trav_enter_instance(const float4 *__restrict__ inst_leaves, float3 world_o, float3 world_d, TravRay &tr, uint2 &ngroup,
                                 uint2 &tgroup, uint32_t &cur_inst, Stack &stack) {
    const int ti = bfind32(tgroup.y);
    tgroup.y &= ~(1u << ti);
    const float4 *rp = inst_leaves + (size_t)(tgroup.x + (uint32_t)ti) * 4;
    const float4 r0 = SOLB_LDG4(rp + 0), r1 = SOLB_LDG4(rp + 1), r2 = SOLB_LDG4(rp + 2), r3 = SOLB_LDG4(rp + 3);
    if (tgroup.y) stack.push(tgroup);                 // the leaf's other instances
    if (ngroup.y & 0xff000000u) stack.push(ngroup);   // TLAS siblings still to visit
    stack.push(SOLB_STACK_SENTINEL);
    const float3 o = f3(fmaf(r0.x, world_o.x, fmaf(r0.y, world_o.y, fmaf(r0.z, world_o.z, r0.w))),
                        fmaf(r1.x, world_o.x, fmaf(r1.y, world_o.y, fmaf(r1.z, world_o.z, r1.w))),
                        fmaf(r2.x, world_o.x, fmaf(r2.y, world_o.y, fmaf(r2.z, world_o.z, r2.w))));
    const float3 d = f3(fmaf(r0.x, world_d.x, fmaf(r0.y, world_d.y, r0.z * world_d.z)),
                        fmaf(r1.x, world_d.x, fmaf(r1.y, world_d.y, r1.z * world_d.z)),
                        fmaf(r2.x, world_d.x, fmaf(r2.y, world_d.y, r2.z * world_d.z)));
    tr = make_trav_ray(o, d, tr.tmin);
    cur_inst = f2u(r3.y);
    ngroup = make_uint2(f2u(r3.x), 0x80000000u);  // "child 7 ^ oct_inv of a virtual parent" whose child_base is the BLAS root
    tgroup = make_uint2(0u, 0u);
}

// Closest-hit traversal, one thread per ray (debug / AO / megakernel / host emulation).
// Stack: push(uint2), pop() -> uint2, empty().
template <bool STATS, class Stack>
SOLB_HD void trace_closest(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris, const Ray &ray, Hit &hit,
                           Stack &stack, TraceCounters *ctr) {
    hit.inst = SOLB_MISS; hit.prim = SOLB_MISS; hit.gtri = SOLB_MISS;
    hit.t = ray.tmax; hit.u = 0.0f; hit.v = 0.0f;
    const TravRay tr = make_trav_ray(ray.o, ray.d, ray.tmin);
    float tmax = ray.tmax;
    uint2 ngroup = SOLB_ROOT_GROUP;
    uint2 tgroup = make_uint2(0u, 0u);
    for (;;) {
        // invariant: ngroup always carries at least one internal-child hit here (only such groups are pushed)
        trav_node_step(nodes, tr, tmax, ngroup, tgroup, stack);
        if (STATS) ctr->nodes++;
        while (tgroup.y) {
            trav_tri_step(tris, tr, tmax, tgroup, hit);
            if (STATS) ctr->tris++;
        }
        if (!(ngroup.y & 0xff000000u)) {
            if (stack.empty()) break;
            ngroup = stack.pop();
        }
    }
}

// Two-level closest hit, one thread per ray.  A plain state machine (triangle / instance step, node step, pop):
// TLAS groups are postponed on the stack while a BLAS is walked.
template <bool STATS, class Stack>
SOLB_HD void trace_closest_2l(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris, const float4 *__restrict__ inst_leaves,
                              const Ray &ray, Hit &hit, Stack &stack, TraceCounters *ctr) {
    hit.inst = SOLB_MISS; hit.prim = SOLB_MISS; hit.gtri = SOLB_MISS;
    hit.t = ray.tmax; hit.u = 0.0f; hit.v = 0.0f;
    TravRay tr = make_trav_ray(ray.o, ray.d, ray.tmin);
    float tmax = ray.tmax;
    uint2 ngroup = SOLB_ROOT_GROUP;
    uint2 tgroup = make_uint2(0u, 0u);
    bool in_blas = false;
    uint32_t cur_inst = SOLB_MISS;
    for (;;) {
        if (tgroup.y) {
            if (in_blas) {
                if (trav_tri_step(tris, tr, tmax, tgroup, hit)) hit.inst = cur_inst;
                if (STATS) ctr->tris++;
            } else {
                trav_enter_instance(inst_leaves, ray.o, ray.d, tr, ngroup, tgroup, cur_inst, stack);
                in_blas = true;
            }
        } else if (ngroup.y & 0xff000000u) {
            trav_node_step(nodes, tr, tmax, ngroup, tgroup, stack);
            if (STATS) ctr->nodes++;
        } else {
            if (stack.empty()) break;
            const uint2 e = stack.pop();
            if (e.y == 0u) {  // sentinel: leave the instance
                tr = make_trav_ray(ray.o, ray.d, ray.tmin);
                in_blas = false;
            } else if (e.y & 0xff000000u) ngroup = e;
            else tgroup = e;
        }
    }
}

// ---- build-time node encoding (used by the collapse kernel; host-testable) -------------------------
struct ChildRef {
    float3 lo, hi;
    uint32_t is_inner;   // 1: internal child, 0: leaf
    uint32_t tri_offset; // leaf: first triangle relative to tri_base
    uint32_t tri_count;  // leaf: 1..SOLB_MAX_LEAF_TRIS
    uint32_t valid;
};

// exponent e (biased) such that 255 * 2^(e-127) >= extent, chosen conservatively
SOLB_HD uint32_t quant_exponent(float extent) {
    if (!(extent > 0.0f)) return 1u;  // flat axis: any tiny positive scale works (q stays 0)
    // smallest power of two s with extent / s <= 255
    float s = extent / 255.0f;
    uint32_t bits = f2u(s);
    uint32_t e = (bits >> 23) & 0xffu;
    if (bits & 0x7fffffu) e += 1;  // round mantissa up to the next power of two
    if (e < 1u) e = 1u;
    if (e > 254u) e = 254u;
    return e;
}

// children[slot] for slot 0..7 (valid == 0: empty).  Internal children must be numbered by the
// caller in slot order starting at child_base.
SOLB_HD void encode_node8(Node8 &out, float3 lo, float3 hi, uint32_t child_base, uint32_t tri_base, const ChildRef *children) {
    uint32_t ex = quant_exponent(hi.x - lo.x), ey = quant_exponent(hi.y - lo.y), ez = quant_exponent(hi.z - lo.z);
    // guard against rounding in (c - lo) / scale: bump the exponent until every child fits in [0, 255]
    for (int axis = 0; axis < 3; axis++) {
        uint32_t &e = axis == 0 ? ex : (axis == 1 ? ey : ez);
        float l = axis == 0 ? lo.x : (axis == 1 ? lo.y : lo.z);
        for (;;) {
            float inv = 1.0f / u2f(e << 23);
            bool ok = true;
            for (int i = 0; i < 8; i++) {
                if (!children[i].valid) continue;
                float ch = axis == 0 ? children[i].hi.x : (axis == 1 ? children[i].hi.y : children[i].hi.z);
                if (ceilf((ch - l) * inv) > 255.0f) ok = false;
            }
            if (ok || e >= 254u) break;
            e++;
        }
    }
    const float sx = u2f(ex << 23), sy = u2f(ey << 23), sz = u2f(ez << 23);
    uint32_t w[20];
    for (int i = 0; i < 20; i++) w[i] = 0;
    w[0] = f2u(lo.x); w[1] = f2u(lo.y); w[2] = f2u(lo.z);
    uint32_t imask = 0;
    uint32_t count0 = 0;  // triangle bits owned by the leaves of slots 0-3
    for (int i = 0; i < 4; i++)
        if (children[i].valid && !children[i].is_inner) count0 += children[i].tri_count;
    for (int i = 0; i < 8; i++) {
        const ChildRef &c = children[i];
        if (!c.valid) continue;
        uint32_t meta = 0;  // tmask byte
        if (c.is_inner) {
            imask |= 1u << i;
        } else {
            const uint32_t unary = c.tri_count == 1 ? 1u : 3u;  // tri_count <= SOLB_MAX_LEAF_TRIS = 2
            meta = unary << (c.tri_offset - (i < 4 ? 0u : count0));
        }
        uint32_t q[6];
        const float cl[3] = { c.lo.x, c.lo.y, c.lo.z }, chh[3] = { c.hi.x, c.hi.y, c.hi.z };
        const float l3[3] = { lo.x, lo.y, lo.z }, s3[3] = { sx, sy, sz };
        for (int a = 0; a < 3; a++) {
            float ql = floorf((cl[a] - l3[a]) / s3[a]);
            float qh = ceilf((chh[a] - l3[a]) / s3[a]);
            ql = fminf(fmaxf(ql, 0.0f), 255.0f);
            qh = fminf(fmaxf(qh, 0.0f), 255.0f);
            // make the decoded box provably contain the child despite rounding
            while (ql > 0.0f && l3[a] + ql * s3[a] > cl[a]) ql -= 1.0f;
            while (qh < 255.0f && l3[a] + qh * s3[a] < chh[a]) qh += 1.0f;
            q[a] = (uint32_t)ql;
            q[3 + a] = (uint32_t)qh;
        }
        const int word = i >> 2, sh = 8 * (i & 3);
        w[6 + word] |= meta << sh;
        w[8 + word] |= q[0] << sh;   // qlo_x
        w[10 + word] |= q[1] << sh;  // qlo_y
        w[12 + word] |= q[2] << sh;  // qlo_z
        w[14 + word] |= q[3] << sh;  // qhi_x
        w[16 + word] |= q[4] << sh;  // qhi_y
        w[18 + word] |= q[5] << sh;  // qhi_z
    }
    w[3] = ex | (ey << 8) | (ez << 16) | (imask << 24);
    w[4] = child_base;
    w[5] = (tri_base & SOLB_TRI_BASE_MASK) | (count0 << 28);
    for (int i = 0; i < 5; i++) out.q[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

// decode what slot i of a node holds (tests / tools): 0 = empty, 1 = internal child, 2 = leaf with tri_count triangles starting at
// triangle tri_offset (relative to the node's triangle base)
SOLB_HD int decode_child_kind(const Node8 &n, int i, uint32_t &tri_offset, uint32_t &tri_count) {
    const uint32_t *w = (const uint32_t *)&n;
    tri_offset = tri_count = 0;
    if ((w[3] >> (24 + i)) & 1u) return 1;
    const uint32_t tmask = (w[6 + (i >> 2)] >> (8 * (i & 3))) & 0xffu;
    if (!tmask) return 0;
    uint32_t first = 0;
    while (!((tmask >> first) & 1u)) first++;
    tri_count = (uint32_t)popc32(tmask);
    tri_offset = first + (i < 4 ? 0u : w[5] >> 28);
    return 2;
}

// decode child box i of a node (tests / invariants)
SOLB_HD void decode_child_box(const Node8 &n, int i, float3 &lo, float3 &hi) {
    const uint32_t *w = (const uint32_t *)&n;
    const uint32_t e = w[3];
    const float sx = u2f((e & 0xffu) << 23), sy = u2f(((e >> 8) & 0xffu) << 23), sz = u2f(((e >> 16) & 0xffu) << 23);
    const int word = i >> 2, sh = 8 * (i & 3);
    lo = f3(u2f(w[0]) + (float)((w[8 + word] >> sh) & 0xff) * sx, u2f(w[1]) + (float)((w[10 + word] >> sh) & 0xff) * sy,
            u2f(w[2]) + (float)((w[12 + word] >> sh) & 0xff) * sz);
    hi = f3(u2f(w[0]) + (float)((w[14 + word] >> sh) & 0xff) * sx, u2f(w[1]) + (float)((w[16 + word] >> sh) & 0xff) * sy,
            u2f(w[2]) + (float)((w[18 + word] >> sh) & 0xff) * sz);
}

}  // namespace solb
