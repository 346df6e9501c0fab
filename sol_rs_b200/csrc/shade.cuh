// shade.cuh — the reference's GLSL stages as per-ray functions (f32, same evaluation order).
//   sampling.glsl        -> tea / next_rand / fresnel / GGX / cosine sampling
//   pathtrace.rgen       -> primary_ray, FrameConsts, resolve_pixel
//   pathtrace.rchit      -> shade_hit
//   pathtrace.rmiss      -> shade_miss
//   ao.rchit / debug.*   -> ao_hit / debug colours
// All SOLB_HD: kernels in trace.cu call them; tests/emu steps the same code on the CPU.
#pragma once
#include "common.cuh"
#include "solb_internal.h"

namespace solb {

// assets/glsl/sampling.glsl:18-32
SOLB_HD uint32_t tea(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}

// assets/glsl/sampling.glsl:35-42.  4294967295.0f == 2^32 in f32, so the division is an exact scale.
SOLB_HD float next_rand(uint32_t &rng) {
    rng = rng * 747796405u + 1u;
    uint32_t word = ((rng >> ((rng >> 28) + 4u)) ^ rng) * 277803737u;
    word = (word >> 22) ^ word;
    return mul_rn((float)word, 2.3283064365386963e-10f);  // u32->f32 round-to-nearest, then * 2^-32 (exact)
}

// assets/glsl/sampling.glsl:52-64
SOLB_HD float fresnel_dielectric(float3 i, float3 m, float eta) {
    float result = 1.0f;
    const float cosThetaI = fabsf(dot(i, m));
    const float sinThetaOSquared = (eta * eta) * (1.0f - cosThetaI * cosThetaI);
    if (sinThetaOSquared <= 1.0f) {
        const float cosThetaO = sh_sqrt(saturate(1.0f - sinThetaOSquared));
        const float Rs = sh_div(cosThetaI - eta * cosThetaO, cosThetaI + eta * cosThetaO);
        const float Rp = sh_div(eta * cosThetaI - cosThetaO, eta * cosThetaI + cosThetaO);
        result = 0.5f * (Rs * Rs + Rp * Rp);
    }
    return result;
}

// assets/glsl/sampling.glsl:66-84
SOLB_HD float3 align_to_direction(float3 n, float cosTheta, float phi) {
    const float sinTheta = sh_sqrt(saturate(1.0f - cosTheta * cosTheta));
    const float s = (n.z < 0.0f ? -1.0f : 1.0f);
    const float a = sh_div(-1.0f, s + n.z);
    const float b = n.x * n.y * a;
    const float3 u = f3(1.0f + s * n.x * n.x * a, s * b, -s * n.x);
    const float3 v = f3(b, s + n.y * n.y * a, -n.y);
    float sp, cp;
#if defined(__CUDA_ARCH__)
    __sincosf(phi, &sp, &cp);  // phi in [0, 2 pi]: absolute error ~2^-21 (GLSL allows 2^-11)
#else
    sp = sinf(phi); cp = cosf(phi);
#endif
    return (u * cp + v * sp) * sinTheta + n * cosTheta;
}

#define SOLB_TWO_PI 6.28318530718f  // sampling.glsl:5

// assets/glsl/sampling.glsl:86-90
SOLB_HD float3 sample_ggx(float3 n, float xi_x, float xi_y, float alphaSquared) {
    const float cosTheta = sh_sqrt(saturate(sh_div(1.0f - xi_x, xi_x * (alphaSquared - 1.0f) + 1.0f)));
    return align_to_direction(n, cosTheta, xi_y * SOLB_TWO_PI);
}
// assets/glsl/sampling.glsl:92-96
SOLB_HD float3 sample_cosine(float3 n, float xi_x, float xi_y) {
    return align_to_direction(n, sh_sqrt(xi_x), xi_y * SOLB_TWO_PI);
}
// GLSL reflect(I, N) = I - 2 dot(N, I) N
SOLB_HD float3 reflect(float3 i, float3 n) { return i - n * (2.0f * dot(n, i)); }
SOLB_HD float3 mix3(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }
SOLB_HD float smoothstep(float e0, float e1, float x) {
    const float t = saturate(sh_div(x - e0, e1 - e0));
    return t * t * (3.0f - 2.0f * t);
}

// Primary ray direction through pixel position (px, py): pathtrace.rgen:52-58, ao.rgen:49-55, debug.rgen:20-26
SOLB_HD float3 primary_dir(const FrameConsts &fc, float px, float py) {
    // (IEEE division and normalisation here, unlike the BRDF arithmetic below: the primary-hit parity gate compares ids bit for bit)
    const float ux = px / (float)fc.width, uy = py / (float)fc.height;
    const float dx = ux * 2.0f - 1.0f, dy = uy * 2.0f - 1.0f;
    const float *P = fc.proj_inv;
    // target = projection_inverse * vec4(d.x, d.y, 1, 1)
    const float3 target = f3(P[0] * dx + P[4] * dy + P[8] + P[12], P[1] * dx + P[5] * dy + P[9] + P[13],
                             P[2] * dx + P[6] * dy + P[10] + P[14]);
    return mat4_mul_dir(fc.view_inv, normalize_ieee(target));
}

struct ShadeVerts {
    float3 p[3], n[3], c[3];
};

SOLB_HD void unpack_shade_record(const float4 *q, ShadeVerts &sv) {
    float f[28];
#pragma unroll
    for (int i = 0; i < 7; i++) {
#if defined(__CUDA_ARCH__)
        const float4 v = __ldg(q + i);
#else
        const float4 v = q[i];
#endif
        f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        sv.p[k] = f3(f[9 * k], f[9 * k + 1], f[9 * k + 2]);
        sv.n[k] = f3(f[9 * k + 3], f[9 * k + 4], f[9 * k + 5]);
        sv.c[k] = f3(f[9 * k + 6], f[9 * k + 7], f[9 * k + 8]);
    }
}

SOLB_HD float3 bary_mix(const float3 *v, float bx, float by, float bz) { return v[0] * bx + v[1] * by + v[2] * bz; }

// ---- base-colour texture (extension; the CPU checker restates it operation for operation) ----
SOLB_HD int tex_wrap(int i, int n, uint32_t mode) {
    if (mode == 33071u) return i < 0 ? 0 : (i >= n ? n - 1 : i);  // CLAMP_TO_EDGE
    if (mode == 33648u) {                                          // MIRRORED_REPEAT
        int m = i % (2 * n);
        if (m < 0) m += 2 * n;
        return m < n ? m : 2 * n - 1 - m;
    }
    int m = i % n;  // REPEAT (glTF default)
    return m < 0 ? m + n : m;
}
SOLB_HD float3 tex_lerp3(float3 a, float3 b, float t) {
    return f3(fmaf(b.x - a.x, t, a.x), fmaf(b.y - a.y, t, a.y), fmaf(b.z - a.z, t, a.z));
}
SOLB_HD float3 tex_fetch(const TexDesc &td, int x, int y) {
#if defined(__CUDA_ARCH__)
    const float4 t = __ldg(td.texels + (size_t)y * td.width + (size_t)x);
#else
    const float4 t = td.texels[(size_t)y * td.width + (size_t)x];
#endif
    return f3(t.x, t.y, t.z);
}
// bilinear, texel centres at (i + 0.5) / n; non-finite coordinates sample (0, 0)
SOLB_HD float3 sample_texture(const TexDesc &td, float u, float v) {
    if (!(fabsf(u) < 1e9f)) u = 0.0f;
    if (!(fabsf(v) < 1e9f)) v = 0.0f;
    const float x = fmaf(u, (float)td.width, -0.5f), y = fmaf(v, (float)td.height, -0.5f);
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int ix0 = tex_wrap((int)x0, (int)td.width, td.wrap_s), ix1 = tex_wrap((int)x0 + 1, (int)td.width, td.wrap_s);
    const int iy0 = tex_wrap((int)y0, (int)td.height, td.wrap_t), iy1 = tex_wrap((int)y0 + 1, (int)td.height, td.wrap_t);
    const float3 top = tex_lerp3(tex_fetch(td, ix0, iy0), tex_fetch(td, ix1, iy0), fx);
    const float3 bot = tex_lerp3(tex_fetch(td, ix0, iy1), tex_fetch(td, ix1, iy1), fx);
    return tex_lerp3(top, bot, fy);
}
// the instance's texture at the hit: uv through indices[] / vertices[] as pathtrace.rchit:61-67 fetches its vertices
SOLB_HD float3 instance_texture(const TexBinding &tb, const DeviceInstance &in, uint32_t prim, float bx, float by, float bz) {
    const uint32_t t1 = f2u(in.mat[10]);
    if (t1 == 0u || t1 > tb.n_tex) return f3(1.0f, 1.0f, 1.0f);
    const uint32_t *ip = tb.indices + in.first_index + 3u * prim;
    float2 uv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float4 q = tb.vertices[(size_t)(in.first_vertex + ip[k]) * 4 + 3];
        uv[k] = make_float2(q.x, q.y);
    }
    const float u = uv[0].x * bx + uv[1].x * by + uv[2].x * bz, v = uv[0].y * bx + uv[1].y * by + uv[2].y * bz;
    return sample_texture(tb.tex[t1 - 1u], u, v);
}

// assets/glsl/pathtrace.rchit:56-113.  Returns true when the path terminates (prd.done = 1).
// inout: origin/dir (ray that hit, replaced by the bounce ray), rng.  out: hit_value.
SOLB_HD bool shade_hit(const DeviceInstance *__restrict__ instances, const ShadeRecord *__restrict__ shade, const TexBinding &texb,
                       uint32_t inst, uint32_t gtri, float u, float v, float3 &origin, float3 &dir, uint32_t &rng, float3 &hit_value) {
    const DeviceInstance &in = instances[inst];
    const float *mat = in.mat;
    if (mat[4] >= 1.0f || mat[5] >= 1.0f || mat[6] >= 1.0f) {  // :71-76 emissive terminates, no RNG draw
        hit_value = f3(mat[4], mat[5], mat[6]);
        return true;
    }
    ShadeVerts sv;
    unpack_shade_record(shade[gtri].q, sv);
    const float bx = 1.0f - u - v, by = u, bz = v;  // :78
    float3 normal = bary_mix(sv.n, bx, by, bz);      // :80
    normal = normalize(mat4_mul_dir(in.transform_it, normal));  // :82
    float3 world_pos = bary_mix(sv.p, bx, by, bz);   // :84
    world_pos = mat4_mul_point(in.transform, world_pos);        // :86
    float3 vertex_color = bary_mix(sv.c, bx, by, bz);           // :88
    if (texb.n_tex) vertex_color = vertex_color * instance_texture(texb, in, gtri - in.shade_first_tri, bx, by, bz);  // (extension)
    const float3 wI = normalize(dir);                           // :90
    const float3 nO = normal * signf_glsl(dot(normal, -wI));    // :91
    const float alphaSquared = mat[9] * mat[9];                 // :92
    const float xi_x = next_rand(rng);                          // :93 (x first, then y)
    const float xi_y = next_rand(rng);
    const float rnd = next_rand(rng);                           // :94
    const float3 base = f3(mat[0], mat[1], mat[2]);
    const float3 in_dir = dir;  // gl_WorldRayDirectionEXT (not normalised)
    origin = world_pos + nO * 0.0001f;  // :96
    if (rnd < mat[8]) {  // :97-100
        dir = sample_ggx(reflect(in_dir, nO), xi_x, xi_y, alphaSquared);
        hit_value = base * vertex_color;
    } else {
        const float3 m = sample_ggx(nO, xi_x, xi_y, alphaSquared);  // :102
        if (rnd < fresnel_dielectric(nO, m, 1.0f / 1.5f)) {         // :103-106
            dir = reflect(in_dir, m);
            hit_value = f3(1.0f, 1.0f, 1.0f);
        } else {  // :107-110
            dir = sample_cosine(nO, xi_x, xi_y);
            hit_value = base * vertex_color;
        }
    }
    return false;
}

// assets/glsl/pathtrace.rmiss:8-21
SOLB_HD float3 shade_miss(uint32_t enable_sky, float3 dir) {
    if (!enable_sky) return f3(0.0f, 0.0f, 0.0f);
    const float3 wI = normalize(dir);
    const float t = smoothstep(0.35f, 0.65f, 0.5f * (wI.y + 1.0f));
    const float3 sky = mix3(f3(0.58f, 0.45f, 0.25f), f3(0.3f, 0.4f, 0.5f), t);
    const bool is_sun = dot(wI, normalize(f3(0.0f, 1.0f, -0.25f))) > 0.99f;
    return mix3(sky, f3(120.0f, 100.0f, 50.0f), is_sun ? 1.0f : 0.0f);
}

// imageStore to rgba8 UNORM: clamp, scale, round-to-nearest-even; NaN -> 0
SOLB_HD uint32_t to_unorm8(float c) {
    if (!(c == c)) return 0u;
    const float x = saturate(c) * 255.0f;
#if defined(__CUDA_ARCH__)
    return (uint32_t)__float2int_rn(x);
#else
    return (uint32_t)nearbyintf(x);
#endif
}
SOLB_HD uint32_t pack_rgba8(float r, float g, float b, float a) {
    return to_unorm8(r) | (to_unorm8(g) << 8) | (to_unorm8(b) << 16) | (to_unorm8(a) << 24);
}
SOLB_HD bool is_nan_or_inf3(float3 c, bool want_nan) {
    const bool n = !(c.x == c.x) || !(c.y == c.y) || !(c.z == c.z);
    if (want_nan) return n;
    return fabsf(c.x) == INFINITY || fabsf(c.y) == INFINITY || fabsf(c.z) == INFINITY;
}

// pathtrace.rgen:88-103.  pixel_sum = sum of the spp sample colours of this frame.
// Returns the new accumulation value and the packed rgba8 display value.
SOLB_HD float4 resolve_pixel(const FrameConsts &fc, float3 pixel_sum, float4 old4, uint32_t &rgba8_out) {
    float3 pixel = pixel_sum * sh_div(1.0f, (float)fc.spp);  // :88
    float4 out;
    if (fc.accum_mode == 1u) {
        // SOLB_ACCUM_SUM (multi-GPU): keep a per-rank sum of frame colours and a frame count;
        // a NaN/Inf frame contributes the current mean (the reference keeps oldColor, :94-99)
        const float3 old = f3(old4.x, old4.y, old4.z);
        float3 add = pixel;
        if (is_nan_or_inf3(pixel, true) || is_nan_or_inf3(pixel, false)) add = old4.w > 0.0f ? old * (1.0f / old4.w) : f3(0, 0, 0);
        out = make_float4(old.x + add.x, old.y + add.y, old.z + add.z, old4.w + 1.0f);
        const float inv = sh_div(1.0f, out.w);
        pixel = f3(out.x * inv, out.y * inv, out.z * inv);
    } else {
        const float alpha = sh_div(1.0f, (float)(uint32_t)(fc.frame + 1u - (uint32_t)fc.accum_start));  // :89
        const float3 old = f3(old4.x, old4.y, old4.z);
        pixel = mix3(old, pixel, alpha);                       // :91
        if (is_nan_or_inf3(pixel, true)) pixel = old;          // :93-95
        if (is_nan_or_inf3(pixel, false)) pixel = old;         // :96-98
        out = make_float4(pixel.x, pixel.y, pixel.z, 1.0f);    // :100
    }
    const float g = 1.0f / 2.2f;  // postprocess.glsl:38-41
    rgba8_out = pack_rgba8(sh_pow(pixel.x, g), sh_pow(pixel.y, g), sh_pow(pixel.z, g), 1.0f);  // :102-103
    return out;
}

// assets/glsl/ao.rchit:53-88.  depth/sample_id as in the payload; blue = rgba8 texels (flipped rows).
SOLB_HD void ao_hit(const DeviceInstance *__restrict__ instances, const ShadeRecord *__restrict__ shade, uint32_t inst,
                    uint32_t gtri, float u, float v, uint32_t px, uint32_t py, const uint32_t *__restrict__ blue, uint32_t blue_w,
                    uint32_t blue_h, uint32_t depth, uint32_t sample_id, float3 &origin, float3 &dir, uint32_t &rng) {
    const DeviceInstance &in = instances[inst];
    ShadeVerts sv;
    unpack_shade_record(shade[gtri].q, sv);
    const float bx = 1.0f - u - v, by = u, bz = v;
    float3 normal = bary_mix(sv.n, bx, by, bz);
    normal = normalize(mat4_mul_dir(in.transform_it, normal));
    float3 world_pos = bary_mix(sv.p, bx, by, bz);
    world_pos = mat4_mul_point(in.transform, world_pos);
    const float3 rd = dir;
    origin = world_pos + rd * 0.00001f;  // ao.rchit:79
    // getBlueRand2(depth + depth * sampleId): ao.rchit:45-51,80
    const uint32_t bi = depth + depth * sample_id;
    const float rx = next_rand(rng), ry = next_rand(rng);
    float fx = (float)px + rx * (float)blue_w, fy = (float)py + ry * (float)blue_h;
    fx = fx - (float)blue_w * floorf(fx / (float)blue_w);  // GLSL mod
    fy = fy - (float)blue_h * floorf(fy / (float)blue_h);
    int cx = (int)fx, cy = (int)fy;
    cx = cx < 0 ? 0 : (cx >= (int)blue_w ? (int)blue_w - 1 : cx);
    cy = cy < 0 ? 0 : (cy >= (int)blue_h ? (int)blue_h - 1 : cy);
    const uint32_t texel = blue[(uint32_t)cy * blue_w + (uint32_t)cx];
    const float xi_x = (float)((texel >> (8 * (bi & 3u))) & 0xffu) / 255.0f;
    const float xi_y = (float)((texel >> (8 * ((bi + 1u) & 3u))) & 0xffu) / 255.0f;
    const float3 hit_norm = normal * signf_glsl(dot(-rd, normal));  // :81
    dir = sample_cosine(hit_norm, xi_x, xi_y);                      // :82
}

}  // namespace solb
