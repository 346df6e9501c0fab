/*
 * oracle.c — CPU restatement of the sol-rs ray-tracing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA library in sol_rs_b200/csrc.  It is never
 * linked into, loaded by, or called from the product path (libsolb.so / sol_rs_b200); only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the reference (num3ric/sol-rs) ships no tests, no golden images and no
 * known-answer vectors, and it cannot be built or run here (no rustc, no Vulkan, no lavapipe).
 * Its BVH build / traversal / ray-triangle test live in the Vulkan driver and have no source.
 * This oracle therefore restates (a) the in-tree GLSL arithmetic line by line in f32 and
 * (b) the driver's contract ("closest hit, opaque, two-sided, tmin < t < tmax") with an f64
 * Moller-Trumbore closest hit over a plain CPU BVH.  The only pins are the integer RNG KATs of
 * SURVEY.md 8c (derived from sampling.glsl by exact arithmetic) checked in tests/test_oracle.py.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference
 * root).  Compile with -ffp-contract=off so no FMA contraction sneaks into the f32 restatement.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * sampling.glsl
 * ---------------------------------------------------------------------------------------- */

/* assets/glsl/sampling.glsl:18-32  tea(): 16 rounds, returns v0 */
ORC_API uint32_t orc_tea(uint32_t val0, uint32_t val1)
{
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}

/* assets/glsl/sampling.glsl:35-42  nextRand(): LCG step + PCG RXS-M-XS output word */
ORC_API uint32_t orc_next_word(uint32_t *rng)
{
    *rng = *rng * 747796405u + 1u;
    uint32_t r = *rng;
    uint32_t word = ((r >> ((r >> 28) + 4u)) ^ r) * 277803737u;
    word = (word >> 22) ^ word;
    return word;
}

/* assets/glsl/sampling.glsl:41  float(word) / 4294967295.0f.  The literal rounds to 2^32 in
 * f32, so the quotient is float(word) * 2^-32 exactly (u32->f32 is round-to-nearest-even);
 * 1.0f is reachable. */
ORC_API float orc_next_rand(uint32_t *rng)
{
    uint32_t word = orc_next_word(rng);
    return (float)word / 4294967295.0f;
}

typedef struct { float x, y, z; } v3;

static inline v3 V3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_scale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float v3_length(v3 a) { return sqrtf(v3_dot(a, a)); }
/* GLSL normalize(v) = v / length(v) */
static inline v3 v3_normalize(v3 a) { float l = v3_length(a); return V3(a.x / l, a.y / l, a.z / l); }
/* GLSL reflect(I, N) = I - 2 * dot(N, I) * N */
static inline v3 v3_reflect(v3 i, v3 n) { float d = 2.0f * v3_dot(n, i); return v3_sub(i, v3_scale(n, d)); }
static inline float saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); } /* sampling.glsl:7 */
static inline float signf(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

/* assets/glsl/sampling.glsl:52-64 */
static float fresnel_dielectric(v3 i, v3 m, float eta)
{
    float result = 1.0f;
    float cosThetaI = fabsf(v3_dot(i, m));
    float sinThetaOSquared = (eta * eta) * (1.0f - cosThetaI * cosThetaI);
    if (sinThetaOSquared <= 1.0f) {
        float cosThetaO = sqrtf(saturatef(1.0f - sinThetaOSquared));
        float Rs = (cosThetaI - eta * cosThetaO) / (cosThetaI + eta * cosThetaO);
        float Rp = (eta * cosThetaI - cosThetaO) / (eta * cosThetaI + cosThetaO);
        result = 0.5f * (Rs * Rs + Rp * Rp);
    }
    return result;
}

/* assets/glsl/sampling.glsl:66-73 */
static void build_orthonormal_basis(v3 n, v3 *u, v3 *v)
{
    float s = (n.z < 0.0f ? -1.0f : 1.0f);
    float a = -1.0f / (s + n.z);
    float b = n.x * n.y * a;
    *u = V3(1.0f + s * n.x * n.x * a, s * b, -s * n.x);
    *v = V3(b, s + n.y * n.y * a, -n.y);
}

/* assets/glsl/sampling.glsl:75-84 */
static v3 align_to_direction(v3 n, float cosTheta, float phi)
{
    float sinTheta = sqrtf(saturatef(1.0f - cosTheta * cosTheta));
    v3 u, v;
    build_orthonormal_basis(n, &u, &v);
    v3 t = v3_add(v3_scale(u, cosf(phi)), v3_scale(v, sinf(phi)));
    return v3_add(v3_scale(t, sinTheta), v3_scale(n, cosTheta));
}

#define ORC_TWO_PI 6.28318530718f /* sampling.glsl:5 */

/* assets/glsl/sampling.glsl:86-90 */
static v3 sample_ggx(v3 n, float xi_x, float xi_y, float alphaSquared)
{
    float cosTheta = sqrtf(saturatef((1.0f - xi_x) / (xi_x * (alphaSquared - 1.0f) + 1.0f)));
    return align_to_direction(n, cosTheta, xi_y * ORC_TWO_PI);
}

/* assets/glsl/sampling.glsl:92-96 */
static v3 sample_cosine(v3 n, float xi_x, float xi_y)
{
    float cosTheta = sqrtf(xi_x);
    return align_to_direction(n, cosTheta, xi_y * ORC_TWO_PI);
}

/* exported wrappers so unit tests can poke the samplers directly */
ORC_API float orc_fresnel_dielectric(const float i[3], const float m[3], float eta)
{ return fresnel_dielectric(V3(i[0], i[1], i[2]), V3(m[0], m[1], m[2]), eta); }
ORC_API void orc_sample_ggx(const float n[3], const float xi[2], float a2, float out[3])
{ v3 r = sample_ggx(V3(n[0], n[1], n[2]), xi[0], xi[1], a2); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
ORC_API void orc_sample_cosine(const float n[3], const float xi[2], float out[3])
{ v3 r = sample_cosine(V3(n[0], n[1], n[2]), xi[0], xi[1]); out[0] = r.x; out[1] = r.y; out[2] = r.z; }

/* ------------------------------------------------------------------------------------------
 * Scene (reference byte layouts)
 * ---------------------------------------------------------------------------------------- */

/* One instance == one primitive section == one BLAS: src/ray/mod.rs:78-134 */
typedef struct {
    uint32_t first_vertex;   /* into vertices[] (ModelVertex units): src/scene/mod.rs:55-61 */
    uint32_t n_vertices;
    uint32_t first_index;    /* into indices[]: src/scene/mod.rs:47-53 */
    uint32_t n_indices;      /* primitive_count = n_indices/3: src/ray/acceleration.rs:183 */
    float transform[16];     /* column-major, SceneInstance.transform: src/ray/mod.rs:22 */
    float transform_it[16];  /* inverse().transpose(): src/ray/mod.rs:29,116 */
    float material[12];      /* MaterialInfo 48 B, materials[gl_InstanceID]: src/scene/mod.rs:19-29 */
} OrcInstance;

typedef struct { double v0[3], e1[3], e2[3]; uint32_t inst, prim; } OrcTri;
typedef struct { double lo[3], hi[3]; uint32_t left, right, first, count; } OrcNode;

typedef struct {
    uint32_t n_inst;
    OrcInstance *inst;
    float *vertices;      /* 16 floats per ModelVertex: src/scene/mesh.rs:9-14 */
    uint32_t *indices;    /* section-relative (SURVEY App.A item 2) */
    uint32_t n_tris;
    OrcTri *tris;         /* world space, BVH leaf order */
    OrcNode *nodes;
    uint32_t n_nodes;
    double scale;         /* max |coordinate| of the world-space scene: sizes the interval-end tolerance */
    /* Base-colour textures: an EXTENSION shared with the product, not reference behaviour - the reference
     * reserves SceneInstance::texture_offset (src/ray/mod.rs:20, pathtrace.rchit:31) and samples nothing.  Unbound
     * (n_tex == 0): the shaders above, untouched. */
    uint32_t n_tex;
    struct OrcTexture *tex;
    uint32_t *inst_tex;   /* per instance: texture index or 0xffffffff */
} OrcScene;

typedef struct OrcTexture {
    uint32_t width, height, wrap_s, wrap_t; /* glTF sampler codes: 10497 repeat, 33071 clamp, 33648 mirrored repeat */
    float *texels;                          /* linear-light rgba, rows top first */
} OrcTexture;

static void bvh_bounds(const OrcTri *t, double lo[3], double hi[3])
{
    for (int k = 0; k < 3; k++) {
        double a = t->v0[k], b = t->v0[k] + t->e1[k], c = t->v0[k] + t->e2[k];
        lo[k] = fmin(a, fmin(b, c));
        hi[k] = fmax(a, fmax(b, c));
    }
}

static int g_axis;
static int cmp_centroid(const void *pa, const void *pb)
{
    const OrcTri *a = (const OrcTri *)pa, *b = (const OrcTri *)pb;
    double ca = a->v0[g_axis] * 3.0 + a->e1[g_axis] + a->e2[g_axis];
    double cb = b->v0[g_axis] * 3.0 + b->e1[g_axis] + b->e2[g_axis];
    return (ca < cb) ? -1 : (ca > cb);
}

/* plain median-split binary BVH; only has to be correct and conservative */
static uint32_t bvh_build(OrcScene *s, uint32_t first, uint32_t count)
{
    uint32_t id = s->n_nodes++;
    OrcNode *n = &s->nodes[id];
    for (int k = 0; k < 3; k++) { n->lo[k] = DBL_MAX; n->hi[k] = -DBL_MAX; }
    for (uint32_t i = first; i < first + count; i++) {
        double lo[3], hi[3];
        bvh_bounds(&s->tris[i], lo, hi);
        for (int k = 0; k < 3; k++) { n->lo[k] = fmin(n->lo[k], lo[k]); n->hi[k] = fmax(n->hi[k], hi[k]); }
    }
    /* pad: the f64 slab test below must never cull a true hit */
    for (int k = 0; k < 3; k++) {
        double pad = 1e-9 * (1.0 + fabs(n->lo[k]) + fabs(n->hi[k]));
        n->lo[k] -= pad; n->hi[k] += pad;
    }
    n->first = first; n->count = count; n->left = n->right = 0;
    if (count <= 4) return id;
    int axis = 0;
    double ext = -1;
    for (int k = 0; k < 3; k++) if (n->hi[k] - n->lo[k] > ext) { ext = n->hi[k] - n->lo[k]; axis = k; }
    g_axis = axis;
    qsort(&s->tris[first], count, sizeof(OrcTri), cmp_centroid);
    uint32_t half = count / 2;
    uint32_t l = bvh_build(s, first, half);
    uint32_t r = bvh_build(s, first + half, count - half);
    n = &s->nodes[id];
    n->left = l; n->right = r; n->count = 0;
    return id;
}

/* mat4 (column-major) * vec4, summed in column order like a GLSL compiler would emit */
static inline void mat4_mul_vec4(const float m[16], const float v[4], float out[4])
{
    for (int r = 0; r < 4; r++)
        out[r] = m[0 + r] * v[0] + m[4 + r] * v[1] + m[8 + r] * v[2] + m[12 + r] * v[3];
}

ORC_API OrcScene *orc_scene_create(uint32_t n_inst, const OrcInstance *inst,
                                   const float *vertices, uint32_t n_vertices,
                                   const uint32_t *indices, uint32_t n_indices)
{
    OrcScene *s = (OrcScene *)calloc(1, sizeof(OrcScene));
    s->n_inst = n_inst;
    s->inst = (OrcInstance *)malloc(sizeof(OrcInstance) * (n_inst ? n_inst : 1));
    memcpy(s->inst, inst, sizeof(OrcInstance) * n_inst);
    s->vertices = (float *)malloc(sizeof(float) * 16 * (n_vertices ? n_vertices : 1));
    memcpy(s->vertices, vertices, sizeof(float) * 16 * n_vertices);
    s->indices = (uint32_t *)malloc(sizeof(uint32_t) * (n_indices ? n_indices : 1));
    memcpy(s->indices, indices, sizeof(uint32_t) * n_indices);
    uint32_t nt = 0;
    for (uint32_t i = 0; i < n_inst; i++) nt += inst[i].n_indices / 3;
    s->n_tris = nt;
    s->tris = (OrcTri *)malloc(sizeof(OrcTri) * (nt ? nt : 1));
    uint32_t t = 0;
    for (uint32_t i = 0; i < n_inst; i++) {
        const OrcInstance *in = &inst[i];
        for (uint32_t p = 0; p < in->n_indices / 3; p++, t++) {
            double w[3][3];
            for (int c = 0; c < 3; c++) {
                /* BLAS geometry: positions = first 12 B of each 64 B vertex, first_vertex added
                 * to the (section-relative) index: src/ray/acceleration.rs:160-162,197.
                 * Instance transform = first three rows of the mesh 4x4: acceleration.rs:330-331.
                 * Object->world done in f64 from the f32 inputs (the "exact" geometry). */
                uint32_t vi = in->first_vertex + indices[in->first_index + 3 * p + c];
                const float *pos = &vertices[16 * (size_t)vi];
                for (int r = 0; r < 3; r++)
                    w[c][r] = (double)in->transform[0 + r] * pos[0] + (double)in->transform[4 + r] * pos[1] +
                              (double)in->transform[8 + r] * pos[2] + (double)in->transform[12 + r];
            }
            OrcTri *tr = &s->tris[t];
            for (int r = 0; r < 3; r++) {
                tr->v0[r] = w[0][r];
                tr->e1[r] = w[1][r] - w[0][r];
                tr->e2[r] = w[2][r] - w[0][r];
            }
            tr->inst = i; tr->prim = p;
        }
    }
    s->nodes = (OrcNode *)malloc(sizeof(OrcNode) * (2 * (size_t)nt + 2));
    s->n_nodes = 0;
    s->scale = 1.0;
    for (uint32_t i = 0; i < nt; i++)
        for (int r = 0; r < 3; r++) {
            const OrcTri *tr = &s->tris[i];
            s->scale = fmax(s->scale, fmax(fabs(tr->v0[r]), fmax(fabs(tr->v0[r] + tr->e1[r]), fabs(tr->v0[r] + tr->e2[r]))));
        }
    if (nt) bvh_build(s, 0, nt);
    return s;
}

static void orc_free_textures(OrcScene *s)
{
    for (uint32_t t = 0; t < s->n_tex; t++) free(s->tex[t].texels);
    free(s->tex); free(s->inst_tex);
    s->tex = NULL; s->inst_tex = NULL; s->n_tex = 0;
}

/* rgba8[t]: width x height x 4 bytes, rows top first; colour channels sRGB-decoded (IEC 61966-2-1) when srgb != 0.
 * inst_tex[i]: texture of instance i (the reference's reserved texture_offset) or 0xffffffff. */
ORC_API void orc_scene_set_textures(OrcScene *s, uint32_t n_tex, const uint8_t *const *rgba8, const uint32_t *width,
                                    const uint32_t *height, const uint32_t *wrap_s, const uint32_t *wrap_t, int srgb,
                                    const uint32_t *inst_tex)
{
    orc_free_textures(s);
    if (!n_tex) return;
    float lut[256];
    for (int i = 0; i < 256; i++) {
        const float x = (float)i / 255.0f;
        lut[i] = x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f);
    }
    s->n_tex = n_tex;
    s->tex = (OrcTexture *)calloc(n_tex, sizeof(OrcTexture));
    for (uint32_t t = 0; t < n_tex; t++) {
        OrcTexture *d = &s->tex[t];
        d->width = width[t]; d->height = height[t];
        d->wrap_s = wrap_s[t] ? wrap_s[t] : 10497u; d->wrap_t = wrap_t[t] ? wrap_t[t] : 10497u;
        const size_t n = (size_t)width[t] * height[t];
        d->texels = (float *)malloc(sizeof(float) * 4 * n);
        for (size_t i = 0; i < n; i++) {
            for (int c = 0; c < 3; c++) d->texels[4 * i + c] = srgb ? lut[rgba8[t][4 * i + c]] : (float)rgba8[t][4 * i + c] / 255.0f;
            d->texels[4 * i + 3] = (float)rgba8[t][4 * i + 3] / 255.0f;
        }
    }
    s->inst_tex = (uint32_t *)malloc(sizeof(uint32_t) * (s->n_inst ? s->n_inst : 1));
    memcpy(s->inst_tex, inst_tex, sizeof(uint32_t) * s->n_inst);
}

static int tex_wrap(int i, int n, uint32_t mode)
{
    if (mode == 33071u) return i < 0 ? 0 : (i >= n ? n - 1 : i);
    if (mode == 33648u) {
        int m = i % (2 * n);
        if (m < 0) m += 2 * n;
        return m < n ? m : 2 * n - 1 - m;
    }
    int m = i % n;
    return m < 0 ? m + n : m;
}

/* bilinear, texel centres at (i + 0.5) / n, lerp as fma(b - a, t, a); non-finite coordinates sample (0, 0) */
static void sample_texture(const OrcTexture *td, float u, float v, float out[3])
{
    if (!(fabsf(u) < 1e9f)) u = 0.0f;
    if (!(fabsf(v) < 1e9f)) v = 0.0f;
    const float x = fmaf(u, (float)td->width, -0.5f), y = fmaf(v, (float)td->height, -0.5f);
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int ix0 = tex_wrap((int)x0, (int)td->width, td->wrap_s), ix1 = tex_wrap((int)x0 + 1, (int)td->width, td->wrap_s);
    const int iy0 = tex_wrap((int)y0, (int)td->height, td->wrap_t), iy1 = tex_wrap((int)y0 + 1, (int)td->height, td->wrap_t);
    const float *t00 = td->texels + 4 * ((size_t)iy0 * td->width + ix0), *t10 = td->texels + 4 * ((size_t)iy0 * td->width + ix1);
    const float *t01 = td->texels + 4 * ((size_t)iy1 * td->width + ix0), *t11 = td->texels + 4 * ((size_t)iy1 * td->width + ix1);
    for (int c = 0; c < 3; c++) {
        const float top = fmaf(t10[c] - t00[c], fx, t00[c]), bot = fmaf(t11[c] - t01[c], fx, t01[c]);
        out[c] = fmaf(bot - top, fy, top);
    }
}

ORC_API void orc_sample_texture(const OrcScene *s, uint32_t t, float u, float v, float out[3])
{
    sample_texture(&s->tex[t], u, v, out);
}

ORC_API void orc_scene_destroy(OrcScene *s)
{
    if (!s) return;
    orc_free_textures(s);
    free(s->inst); free(s->vertices); free(s->indices); free(s->tris); free(s->nodes); free(s);
}

ORC_API uint32_t orc_scene_tri_count(const OrcScene *s) { return s->n_tris; }

ORC_API void orc_scene_bounds(const OrcScene *s, double lo[3], double hi[3])
{
    for (int k = 0; k < 3; k++) { lo[k] = DBL_MAX; hi[k] = -DBL_MAX; }
    for (uint32_t i = 0; i < s->n_tris; i++) {
        double l[3], h[3];
        bvh_bounds(&s->tris[i], l, h);
        for (int k = 0; k < 3; k++) { lo[k] = fmin(lo[k], l[k]); hi[k] = fmax(hi[k], h[k]); }
    }
}

/* ------------------------------------------------------------------------------------------
 * traceRayEXT contract (assets/glsl/pathtrace.rgen:65-76; src/ray/acceleration.rs:335-337):
 * closest hit, opaque, cull disabled (two-sided), mask 0xFF, tmin < t < tmax.
 * ---------------------------------------------------------------------------------------- */

#define ORC_MISS 0xffffffffu
#define ORC_FLAG_EDGE 1u  /* closest hit lies within eps of a triangle edge/vertex          */
#define ORC_FLAG_TIE 2u   /* another triangle is hit within a relative dt of the closest    */
#define ORC_FLAG_NEAR 4u  /* a triangle narrowly missed (bary > -eps) in front of the result */

typedef struct {
    uint32_t inst, prim;
    float u, v;        /* attribs.xy: weights of vertex 1 and vertex 2 */
    double t;
    uint32_t flags;
} OrcHit;

static inline int slab(const OrcNode *n, const double o[3], const double id[3], double tmin, double tmax, double *tnear)
{
    double t0 = tmin, t1 = tmax;
    for (int k = 0; k < 3; k++) {
        double a = (n->lo[k] - o[k]) * id[k], b = (n->hi[k] - o[k]) * id[k];
        if (a != a || b != b) continue; /* 0 * inf: origin on the slab plane of a flat axis */
        double lo = fmin(a, b), hi = fmax(a, b);
        if (lo > t0) t0 = lo;
        if (hi < t1) t1 = hi;
    }
    *tnear = t0;
    return t0 <= t1;
}

/* classify != 0 additionally fills the edge/tie/near flags (primary-ray parity lists) */
static void closest_hit(const OrcScene *s, const float org[3], const float dir[3], float tmin_f, float tmax_f,
                        int classify, double eps_b, double eps_t, OrcHit *out)
{
    out->inst = out->prim = ORC_MISS; out->u = out->v = 0.0f; out->t = 0.0; out->flags = 0;
    if (!s->n_tris) return;
    const double o[3] = { org[0], org[1], org[2] }, d[3] = { dir[0], dir[1], dir[2] };
    const double tmin = tmin_f, tmax = tmax_f;
    double id[3];
    for (int k = 0; k < 3; k++) id[k] = 1.0 / d[k];
    double best_t = tmax, best_u = 0, best_v = 0;
    uint32_t best = ORC_MISS;
    /* classification bookkeeping */
    double second_t = DBL_MAX;      /* nearest strict hit on a different triangle */
    double loose_t = DBL_MAX;       /* nearest "almost hit" (bary in [-eps, 0)) */
    int end_tie = 0;                /* a candidate sits at tmin / tmax within position rounding */
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const OrcNode *n = &s->nodes[stack[--sp]];
        double tn;
        double limit = classify ? fmin(tmax, best_t * (1.0 + 4.0 * eps_t) + 1e-12) : best_t;
        if (!slab(n, o, id, tmin, limit, &tn)) continue;
        if (n->count == 0) {
            /* push far child first */
            double tl, tr;
            int hl = slab(&s->nodes[n->left], o, id, tmin, limit, &tl);
            int hr = slab(&s->nodes[n->right], o, id, tmin, limit, &tr);
            if (hl && hr) {
                if (tl <= tr) { stack[sp++] = n->right; stack[sp++] = n->left; }
                else { stack[sp++] = n->left; stack[sp++] = n->right; }
            } else if (hl) stack[sp++] = n->left;
            else if (hr) stack[sp++] = n->right;
            continue;
        }
        for (uint32_t i = n->first; i < n->first + n->count; i++) {
            const OrcTri *tr = &s->tris[i];
            /* Moller-Trumbore in f64 */
            double p[3] = { d[1] * tr->e2[2] - d[2] * tr->e2[1], d[2] * tr->e2[0] - d[0] * tr->e2[2], d[0] * tr->e2[1] - d[1] * tr->e2[0] };
            double det = tr->e1[0] * p[0] + tr->e1[1] * p[1] + tr->e1[2] * p[2];
            if (det == 0.0) continue;
            double inv = 1.0 / det;
            double tv[3] = { o[0] - tr->v0[0], o[1] - tr->v0[1], o[2] - tr->v0[2] };
            double u = (tv[0] * p[0] + tv[1] * p[1] + tv[2] * p[2]) * inv;
            double q[3] = { tv[1] * tr->e1[2] - tv[2] * tr->e1[1], tv[2] * tr->e1[0] - tv[0] * tr->e1[2], tv[0] * tr->e1[1] - tv[1] * tr->e1[0] };
            double v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv;
            double t = (tr->e2[0] * q[0] + tr->e2[1] * q[1] + tr->e2[2] * q[2]) * inv;
            double w = 1.0 - u - v;
            double mn = fmin(u, fmin(v, w));
            if (classify && mn >= -eps_b) {
                /* Interval-end tie: the plane of this triangle passes within an f32-position tolerance of the
                 * ray's start (o + tmin d) or end point, so whether t > tmin (t < tmax) holds is decided by
                 * rounding.  |t - tmin| * |d . n^| is that distance. */
                double nx = tr->e1[1] * tr->e2[2] - tr->e1[2] * tr->e2[1], ny = tr->e1[2] * tr->e2[0] - tr->e1[0] * tr->e2[2],
                       nz = tr->e1[0] * tr->e2[1] - tr->e1[1] * tr->e2[0];
                double nl = sqrt(nx * nx + ny * ny + nz * nz);
                double dn = nl > 0.0 ? fabs(d[0] * nx + d[1] * ny + d[2] * nz) / nl : 0.0;
                double tol = 2e-6 * s->scale;
                if (fabs(t - tmin) * dn <= tol || fabs(t - tmax) * dn <= tol) end_tie = 1;
            }
            if (!(t > tmin && t < tmax)) continue;
            if (mn >= 0.0) {
                if (t < best_t || (t == best_t && best != ORC_MISS && i < best)) {
                    if (best != ORC_MISS && best_t < second_t) second_t = best_t;
                    best_t = t; best = i; best_u = u; best_v = v;
                } else if (t < second_t) second_t = t;
            } else if (classify && mn >= -eps_b) {
                if (t < loose_t) loose_t = t;
            }
        }
    }
    if (best != ORC_MISS) {
        const OrcTri *tr = &s->tris[best];
        out->inst = tr->inst; out->prim = tr->prim;
        out->u = (float)best_u; out->v = (float)best_v; out->t = best_t;
        if (classify) {
            double w = 1.0 - best_u - best_v;
            if (fmin(best_u, fmin(best_v, w)) < eps_b) out->flags |= ORC_FLAG_EDGE;
            if (second_t <= best_t * (1.0 + eps_t) + 1e-12) out->flags |= ORC_FLAG_TIE;
            if (loose_t <= best_t * (1.0 + eps_t) + 1e-12) out->flags |= ORC_FLAG_NEAR;
        }
    } else if (classify && loose_t < DBL_MAX) {
        out->flags |= ORC_FLAG_NEAR;
    }
    if (classify && end_tie) out->flags |= ORC_FLAG_TIE;
}

/* Batch closest hit for arbitrary rays: rays[i] = {ox,oy,oz,tmin, dx,dy,dz,tmax}.
 * hits[i] = {inst, prim, bits(u), bits(v)}; t_out optional; flags_out optional (classify). */
ORC_API void orc_trace_rays(const OrcScene *s, const float *rays, uint32_t n, uint32_t *hits, float *t_out,
                            uint8_t *flags_out, double eps_b, double eps_t)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float *r = &rays[8 * i];
        OrcHit h;
        closest_hit(s, r, r + 4, r[3], r[7], flags_out != NULL, eps_b, eps_t, &h);
        hits[4 * i + 0] = h.inst; hits[4 * i + 1] = h.prim;
        memcpy(&hits[4 * i + 2], &h.u, 4); memcpy(&hits[4 * i + 3], &h.v, 4);
        if (t_out) t_out[i] = (float)h.t;
        if (flags_out) flags_out[i] = (uint8_t)h.flags;
    }
}

/* ------------------------------------------------------------------------------------------
 * SceneUniforms: examples/5-pathtrace.rs:7-17, assets/glsl/pathtrace.rgen:13-21 (400 B)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float model[16], view[16], view_inverse[16], projection[16], projection_inverse[16], mvp[16];
    uint32_t frame[3];
    uint32_t pad;
} OrcUniforms;

/* Primary ray: pathtrace.rgen:52-58 == ao.rgen:49-55 == debug.rgen:20-26 */
static void primary_ray(const OrcUniforms *u, float px, float py, uint32_t W, uint32_t H, v3 *origin, v3 *direction)
{
    float inUVx = px / (float)W, inUVy = py / (float)H;
    float dx = inUVx * 2.0f - 1.0f, dy = inUVy * 2.0f - 1.0f;
    float o4[4], t4[4], d4[4];
    const float zero1[4] = { 0, 0, 0, 1 };
    mat4_mul_vec4(u->view_inverse, zero1, o4);
    const float tgt[4] = { dx, dy, 1, 1 };
    mat4_mul_vec4(u->projection_inverse, tgt, t4);
    v3 tn = v3_normalize(V3(t4[0], t4[1], t4[2]));
    const float tn4[4] = { tn.x, tn.y, tn.z, 0 };
    mat4_mul_vec4(u->view_inverse, tn4, d4);
    *origin = V3(o4[0], o4[1], o4[2]);
    *direction = V3(d4[0], d4[1], d4[2]);
}

/* imageStore to an rgba8 UNORM image: clamp to [0,1], scale by 255, round to nearest even
 * (Vulkan spec float->unorm); NaN converts to 0. */
static inline uint8_t to_unorm8(float c)
{
    if (!(c == c)) return 0;
    float x = fminf(fmaxf(c, 0.0f), 1.0f) * 255.0f;
    return (uint8_t)nearbyintf(x);
}

/* ------------------------------------------------------------------------------------------
 * 3-ray-debug pass: assets/glsl/debug.rgen:18-37, debug.rchit:9-13, debug.rmiss:6-9
 * Also emits the (instance, primitive) ids the parity gate needs (the shader itself does not).
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_debug(const OrcScene *s, const OrcUniforms *u, uint32_t W, uint32_t H,
                       uint8_t *rgba8, uint32_t *ids, float *bary_t, uint8_t *flags, double eps_b, double eps_t)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++) {
            v3 o, d;
            primary_ray(u, (float)x + 0.5f, (float)y + 0.5f, W, H, &o, &d); /* debug.rgen:20 */
            const float of[3] = { o.x, o.y, o.z }, df[3] = { d.x, d.y, d.z };
            OrcHit h;
            closest_hit(s, of, df, 0.001f, 1000.0f, flags != NULL, eps_b, eps_t, &h); /* debug.rgen:32-34 */
            v3 hv = d; /* debug.rgen:28: payload pre-set to the direction; miss leaves it */
            if (h.inst != ORC_MISS) hv = V3(1.0f - h.u - h.v, h.u, h.v); /* debug.rchit:11-12 */
            size_t p = (size_t)y * W + x;
            if (rgba8) {
                rgba8[4 * p + 0] = to_unorm8(hv.x); rgba8[4 * p + 1] = to_unorm8(hv.y);
                rgba8[4 * p + 2] = to_unorm8(hv.z); rgba8[4 * p + 3] = 0; /* debug.rgen:36 alpha 0.0 */
            }
            if (ids) { ids[2 * p] = h.inst; ids[2 * p + 1] = h.prim; }
            if (bary_t) { bary_t[3 * p] = h.u; bary_t[3 * p + 1] = h.v; bary_t[3 * p + 2] = (float)h.t; }
            if (flags) flags[p] = (uint8_t)h.flags;
        }
}

/* ------------------------------------------------------------------------------------------
 * 5-pathtrace
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    v3 hitValue; uint32_t depth, sampleId, done; v3 rayOrigin, rayDir; float rmin, rmax; uint32_t rng;
} Payload; /* assets/glsl/payload.glsl:4-15 */

/* pathtrace.rgen:28-37 (also ao.rgen:26-35) */
static void prepare_payload(Payload *prd, v3 origin, v3 direction)
{
    prd->hitValue = V3(0, 0, 0);
    prd->depth = 0;
    prd->done = 0;
    prd->rayOrigin = origin;
    prd->rayDir = direction;
    prd->rmin = fmaxf(1.0f, v3_length(origin)) * 1e-3f;
    prd->rmax = 10000.0f;
}

static inline const float *vertex_of(const OrcScene *s, uint32_t inst, uint32_t prim, int c)
{
    const OrcInstance *in = &s->inst[inst];
    /* pathtrace.rchit:61-67: indices[objId].i[3*prim+c], vertices[objId].v[ind]; descriptors start
     * at the section's first index / first vertex (src/scene/mod.rs:47-61) */
    uint32_t vi = in->first_vertex + s->indices[in->first_index + 3 * prim + c];
    return &s->vertices[16 * (size_t)vi];
}

static inline v3 mix3(const float *a, const float *b, const float *c, float bx, float by, float bz)
{
    /* v0 * b.x + v1 * b.y + v2 * b.z, left to right: pathtrace.rchit:80,84,88 */
    return V3(a[0] * bx + b[0] * by + c[0] * bz, a[1] * bx + b[1] * by + c[1] * bz, a[2] * bx + b[2] * by + c[2] * bz);
}

/* assets/glsl/pathtrace.rchit:56-113 */
static void pathtrace_rchit(const OrcScene *s, const OrcHit *h, v3 worldRayDir, Payload *prd)
{
    const OrcInstance *in = &s->inst[h->inst];
    const float *v0 = vertex_of(s, h->inst, h->prim, 0);
    const float *v1 = vertex_of(s, h->inst, h->prim, 1);
    const float *v2 = vertex_of(s, h->inst, h->prim, 2);
    const float *mat = in->material; /* base_color @0, emissive @4, metallic @8, roughness @9 */
    if (mat[4] >= 1.0f || mat[5] >= 1.0f || mat[6] >= 1.0f) { /* :71-76 */
        prd->hitValue = V3(mat[4], mat[5], mat[6]);
        prd->done = 1;
        prd->depth++;
        return;
    }
    float bx = 1.0f - h->u - h->v, by = h->u, bz = h->v; /* :78 */
    v3 normal = mix3(v0 + 8, v1 + 8, v2 + 8, bx, by, bz); /* normal @ float 8 of ModelVertex */
    {
        const float n4[4] = { normal.x, normal.y, normal.z, 0.0f };
        float r[4];
        mat4_mul_vec4(in->transform_it, n4, r); /* :82 */
        normal = v3_normalize(V3(r[0], r[1], r[2]));
    }
    v3 worldPos = mix3(v0, v1, v2, bx, by, bz); /* :84 */
    {
        const float p4[4] = { worldPos.x, worldPos.y, worldPos.z, 1.0f };
        float r[4];
        mat4_mul_vec4(in->transform, p4, r); /* :86 */
        worldPos = V3(r[0], r[1], r[2]);
    }
    v3 vertex_color = mix3(v0 + 4, v1 + 4, v2 + 4, bx, by, bz); /* :88 */
    if (s->n_tex && s->inst_tex[h->inst] < s->n_tex) { /* extension: albedo x base-colour texture at the interpolated uv */
        const float u = v0[12] * bx + v1[12] * by + v2[12] * bz, v = v0[13] * bx + v1[13] * by + v2[13] * bz; /* uv @ float 12 */
        float tc[3];
        sample_texture(&s->tex[s->inst_tex[h->inst]], u, v, tc);
        vertex_color = V3(vertex_color.x * tc[0], vertex_color.y * tc[1], vertex_color.z * tc[2]);
    }
    v3 wI = v3_normalize(worldRayDir); /* :90 */
    v3 nO = v3_scale(normal, signf(v3_dot(normal, V3(-wI.x, -wI.y, -wI.z)))); /* :91 */
    float alphaSquared = mat[9] * mat[9]; /* :92 */
    float xi_x = orc_next_rand(&prd->rng); /* :93 nextRand2: x then y (sampling.glsl:44-47) */
    float xi_y = orc_next_rand(&prd->rng);
    float rnd = orc_next_rand(&prd->rng); /* :94 */
    v3 base = V3(mat[0], mat[1], mat[2]);
    prd->rayOrigin = v3_add(worldPos, v3_scale(nO, 0.0001f)); /* :96 */
    if (rnd < mat[8]) { /* :97-100 */
        prd->rayDir = sample_ggx(v3_reflect(worldRayDir, nO), xi_x, xi_y, alphaSquared);
        prd->hitValue = v3_mul(base, vertex_color);
    } else {
        v3 m = sample_ggx(nO, xi_x, xi_y, alphaSquared); /* :102 */
        if (rnd < fresnel_dielectric(nO, m, 1.0f / 1.5f)) { /* :103-106 */
            prd->rayDir = v3_reflect(worldRayDir, m);
            prd->hitValue = V3(1.0f, 1.0f, 1.0f);
        } else { /* :107-110 */
            prd->rayDir = sample_cosine(nO, xi_x, xi_y);
            prd->hitValue = v3_mul(base, vertex_color);
        }
    }
    prd->depth++; /* :112 */
}

static inline float smoothstepf(float e0, float e1, float x)
{
    float t = saturatef((x - e0) / (e1 - e0));
    return t * t * (3.0f - 2.0f * t);
}
static inline v3 mixv(v3 a, v3 b, float t) /* GLSL mix: a*(1-t) + b*t */
{
    return V3(a.x * (1.0f - t) + b.x * t, a.y * (1.0f - t) + b.y * t, a.z * (1.0f - t) + b.z * t);
}

/* assets/glsl/pathtrace.rmiss:8-21 */
static void pathtrace_rmiss(int enable_sky, v3 worldRayDir, Payload *prd)
{
    if (enable_sky) {
        v3 wI = v3_normalize(worldRayDir);
        float t = smoothstepf(0.35f, 0.65f, 0.5f * (wI.y + 1.0f));
        v3 sky = mixv(V3(0.58f, 0.45f, 0.25f), V3(0.3f, 0.4f, 0.5f), t);
        int isSun = v3_dot(wI, v3_normalize(V3(0.0f, 1.0f, -0.25f))) > 0.99f;
        prd->hitValue = mixv(sky, V3(120.0f, 100.0f, 50.0f), isSun ? 1.0f : 0.0f);
    } else {
        prd->hitValue = V3(0, 0, 0);
    }
    prd->done = 1;
}

typedef struct {
    uint64_t rays;      /* traceRayEXT calls */
    uint64_t hits;      /* closest-hit invocations */
    uint64_t paths;     /* samples started */
    uint64_t capped;    /* paths ended by depth > maxBounces */
    uint64_t emissive;  /* paths ended on an emitter */
} OrcStats;

/* One reference launch of pathtrace.rgen over W x H pixels (assets/glsl/pathtrace.rgen:39-104).
 * accum: float4[W*H] in/out (accumImage, rgba32f); render: rgba8[W*H] out (renderImage).
 * spp / max_bounces are the shader literals 8 / 32 (:43-44) exposed as parameters. */
ORC_API void orc_pathtrace_frame(const OrcScene *s, const OrcUniforms *u, uint32_t W, uint32_t H,
                                 int32_t accum_start_frame, int enable_sky, int spp, int max_bounces,
                                 float *accum, uint8_t *render, OrcStats *stats)
{
    uint64_t n_rays = 0, n_hits = 0, n_paths = 0, n_capped = 0, n_emis = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : n_rays, n_hits, n_paths, n_capped, n_emis)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++) {
            Payload prd;
            memset(&prd, 0, sizeof(prd));
            v3 pixel = V3(0, 0, 0);
            prd.rng = orc_tea(x + (uint32_t)y * W, u->frame[2]); /* :47 */
            for (int i = 0; i < spp; i++) {
                float jx = orc_next_rand(&prd.rng); /* :52 vec2(nextRand, nextRand): left to right */
                float jy = orc_next_rand(&prd.rng);
                v3 o, d;
                primary_ray(u, (float)x + jx, (float)y + jy, W, H, &o, &d); /* :52-58 */
                prepare_payload(&prd, o, d); /* :60 */
                prd.sampleId = (uint32_t)i;
                v3 thr = V3(1.0f, 1.0f, 1.0f);
                n_paths++;
                for (;;) {
                    const float of[3] = { prd.rayOrigin.x, prd.rayOrigin.y, prd.rayOrigin.z };
                    const float df[3] = { prd.rayDir.x, prd.rayDir.y, prd.rayDir.z };
                    OrcHit h;
                    closest_hit(s, of, df, prd.rmin, prd.rmax, 0, 0, 0, &h); /* :65-76 */
                    n_rays++;
                    if (h.inst != ORC_MISS) {
                        n_hits++;
                        pathtrace_rchit(s, &h, prd.rayDir, &prd);
                        if (prd.done) n_emis++;
                    } else
                        pathtrace_rmiss(enable_sky, prd.rayDir, &prd);
                    thr = v3_mul(thr, prd.hitValue); /* :77 */
                    if (prd.done == 1) break;        /* :78-80 */
                    if ((int)prd.depth > max_bounces) { /* :81-84 */
                        thr = V3(0, 0, 0);
                        n_capped++;
                        break;
                    }
                }
                pixel = v3_add(pixel, thr); /* :86 */
            }
            pixel = v3_scale(pixel, 1.0f / (float)spp); /* :88 */
            size_t p = (size_t)y * W + x;
            /* :89-101 */
            float alpha = 1.0f / (float)(uint32_t)(u->frame[2] + 1u - (uint32_t)accum_start_frame);
            v3 old = V3(accum[4 * p], accum[4 * p + 1], accum[4 * p + 2]);
            pixel = mixv(old, pixel, alpha);
            if (isnan(pixel.x) || isnan(pixel.y) || isnan(pixel.z)) pixel = old;
            if (isinf(pixel.x) || isinf(pixel.y) || isinf(pixel.z)) pixel = old;
            accum[4 * p] = pixel.x; accum[4 * p + 1] = pixel.y; accum[4 * p + 2] = pixel.z; accum[4 * p + 3] = 1.0f;
            if (render) {
                /* :102-103 + postprocess.glsl:38-41 */
                render[4 * p + 0] = to_unorm8(powf(pixel.x, 1.0f / 2.2f));
                render[4 * p + 1] = to_unorm8(powf(pixel.y, 1.0f / 2.2f));
                render[4 * p + 2] = to_unorm8(powf(pixel.z, 1.0f / 2.2f));
                render[4 * p + 3] = 255;
            }
        }
    if (stats) {
        stats->rays += n_rays; stats->hits += n_hits; stats->paths += n_paths;
        stats->capped += n_capped; stats->emissive += n_emis;
    }
}

/* ------------------------------------------------------------------------------------------
 * 4-ray-ao: assets/glsl/ao.rgen:37-83, ao.rchit:45-88, ao.rmiss:7-10
 * blue: rgba8[256*256], already flipped vertically like src/texture.rs:491
 * image: float4[W*H] in/out (the example binds an rgba32f image: examples/4-ray-ao.rs:58)
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_ao_frame(const OrcScene *s, const OrcUniforms *u, uint32_t W, uint32_t H,
                          int32_t accum_start_frame, const uint8_t *blue, uint32_t blue_w, uint32_t blue_h,
                          float *image, OrcStats *stats)
{
    uint64_t n_rays = 0, n_hits = 0, n_paths = 0;
    const int max_samples = 4, sample_count = 4; /* ao.rgen:42-43 */
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : n_rays, n_hits, n_paths)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++) {
            Payload prd;
            memset(&prd, 0, sizeof(prd));
            v3 ao = V3(0, 0, 0);
            prd.rng = orc_tea(x + (uint32_t)y * W, u->frame[2]); /* ao.rgen:45 */
            for (int i = 0; i < sample_count; i++) {
                float jx = orc_next_rand(&prd.rng); /* ao.rgen:49 nextRand2 */
                float jy = orc_next_rand(&prd.rng);
                v3 o, d;
                primary_ray(u, (float)x + jx, (float)y + jy, W, H, &o, &d);
                prepare_payload(&prd, o, d);
                prd.sampleId = (uint32_t)i;
                n_paths++;
                for (;;) {
                    const float of[3] = { prd.rayOrigin.x, prd.rayOrigin.y, prd.rayOrigin.z };
                    const float df[3] = { prd.rayDir.x, prd.rayDir.y, prd.rayDir.z };
                    OrcHit h;
                    closest_hit(s, of, df, prd.rmin, prd.rmax, 0, 0, 0, &h);
                    n_rays++;
                    if (h.inst != ORC_MISS) {
                        n_hits++;
                        /* ao.rchit:53-88 */
                        const OrcInstance *in = &s->inst[h.inst];
                        const float *v0 = vertex_of(s, h.inst, h.prim, 0);
                        const float *v1 = vertex_of(s, h.inst, h.prim, 1);
                        const float *v2 = vertex_of(s, h.inst, h.prim, 2);
                        float bx = 1.0f - h.u - h.v, by = h.u, bz = h.v;
                        v3 normal = mix3(v0 + 8, v1 + 8, v2 + 8, bx, by, bz);
                        const float n4[4] = { normal.x, normal.y, normal.z, 0.0f };
                        float r[4];
                        mat4_mul_vec4(in->transform_it, n4, r);
                        normal = v3_normalize(V3(r[0], r[1], r[2]));
                        v3 wp = mix3(v0, v1, v2, bx, by, bz);
                        const float p4[4] = { wp.x, wp.y, wp.z, 1.0f };
                        mat4_mul_vec4(in->transform, p4, r);
                        wp = V3(r[0], r[1], r[2]);
                        v3 rd = prd.rayDir; /* gl_WorldRayDirectionEXT */
                        prd.rayOrigin = v3_add(wp, v3_scale(rd, 0.00001f)); /* ao.rchit:79 */
                        /* getBlueRand2(prd.depth + prd.depth * prd.sampleId): ao.rchit:45-51,80 */
                        uint32_t bi = prd.depth + prd.depth * prd.sampleId;
                        float rx = orc_next_rand(&prd.rng), ry = orc_next_rand(&prd.rng);
                        float fx = (float)x + rx * (float)blue_w, fy = (float)y + ry * (float)blue_h;
                        /* GLSL mod(a, b) = a - b * floor(a / b) */
                        fx = fx - (float)blue_w * floorf(fx / (float)blue_w);
                        fy = fy - (float)blue_h * floorf(fy / (float)blue_h);
                        int cx = (int)fx, cy = (int)fy;
                        if (cx < 0) cx = 0; if (cy < 0) cy = 0;
                        /* texelFetch outside the image is undefined; clamp like robust access would not hit */
                        if (cx >= (int)blue_w) cx = (int)blue_w - 1;
                        if (cy >= (int)blue_h) cy = (int)blue_h - 1;
                        const uint8_t *texel = &blue[4 * ((size_t)cy * blue_w + cx)];
                        float xi_x = (float)texel[bi % 4] / 255.0f, xi_y = (float)texel[(bi + 1) % 4] / 255.0f;
                        v3 hitNorm = v3_scale(normal, signf(v3_dot(V3(-rd.x, -rd.y, -rd.z), normal))); /* :81 */
                        prd.rayDir = sample_cosine(hitNorm, xi_x, xi_y); /* :82 */
                        prd.rmin = 0.001f; prd.rmax = 10.0f; /* :83 */
                        if (prd.depth > 0) prd.hitValue = v3_add(prd.hitValue, V3(1, 1, 1)); /* :84-86 */
                        prd.depth++;
                    } else
                        prd.done = 1; /* ao.rmiss:9 */
                    if (prd.done == 1 || (int)prd.depth > max_samples) break; /* ao.rgen:71-72 */
                }
                ao = v3_add(ao, v3_scale(prd.hitValue, 1.0f / (float)max_samples)); /* ao.rgen:74 */
            }
            float inv = 1.0f / (float)sample_count; /* ao / float(sample_count) */
            v3 color = V3(1.0f - ao.x / (float)sample_count, 1.0f - ao.y / (float)sample_count, 1.0f - ao.z / (float)sample_count);
            (void)inv;
            size_t p = (size_t)y * W + x;
            float a = 1.0f / (float)(uint32_t)(u->frame[2] - (uint32_t)accum_start_frame + 1u); /* ao.rgen:78 */
            v3 old = V3(image[4 * p], image[4 * p + 1], image[4 * p + 2]);
            color = mixv(old, color, a); /* :80 */
            image[4 * p] = color.x; image[4 * p + 1] = color.y; image[4 * p + 2] = color.z; image[4 * p + 3] = 1.0f; /* :82 */
        }
    if (stats) { stats->rays += n_rays; stats->hits += n_hits; stats->paths += n_paths; }
}

/* launchers such as torchrun export OMP_NUM_THREADS=1 to every rank: the timed CPU arms ask for the host's threads explicitly */
ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
