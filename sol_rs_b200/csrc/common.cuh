// common.cuh — shared types and small math for libsolb (sm_100a).
// Device functions that hold algorithmic logic are SOLB_HD so the test-only host emulation
// (tests/emu) can single-step the exact same code on the CPU; the shipped library only ever
// launches them from kernels.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#include <cuda_runtime.h>  // vector types + make_float3/4 (also usable from a plain host compiler)
#if defined(__CUDACC__)
#define SOLB_HD __host__ __device__ __forceinline__
#else
#define SOLB_HD inline
#endif

namespace solb {


SOLB_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
SOLB_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
SOLB_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// index of the highest set bit (x != 0)
SOLB_HD int bfind32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz(x);
#else
    return 31 - __builtin_clz(x);
#endif
}
SOLB_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
SOLB_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

// ---- arithmetic that must never be contracted into FMAs (watertightness, RNG->float) ----
SOLB_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
SOLB_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
SOLB_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}

// fast reciprocal / division for quantities whose last ulp does not matter (slab offsets, ray-space frame,
// hit depth, barycentrics): MUFU.RCP / MUFU.RSQ instead of the IEEE sequences
SOLB_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    return __fdividef(1.0f, x);
#else
    return 1.0f / x;
#endif
}
SOLB_HD float fast_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
SOLB_HD float fast_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

// Shading arithmetic (shade.cuh): the GLSL these functions restate runs under Vulkan's precision rules, not IEEE ones
// (a / b within 2.5 ULP, inversesqrt 2 ULP, sqrt / normalize / pow inherited from those, sin / cos 2^-11 absolute), so on
// the device they use the hardware approximations (MUFU.RCP / RSQ / SQRT / SIN / COS / LG2 / EX2, all within ~2 ULP or 2^-21
// absolute) instead of the IEEE sequences with their slow-path calls: ~40 % fewer instructions in the shade step of the
// persistent kernel and a third less code in its instruction cache.  The host build (tests/emu) keeps libm.
SOLB_HD float sh_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
SOLB_HD float sh_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}
SOLB_HD float sh_pow(float x, float y) {
#if defined(__CUDA_ARCH__)
    return __powf(x, y);
#else
    return powf(x, y);
#endif
}

// ---- float3 helpers ----
SOLB_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
SOLB_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
SOLB_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
SOLB_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
SOLB_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
SOLB_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
SOLB_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
SOLB_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SOLB_HD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
SOLB_HD float length(float3 a) { return sqrtf(dot(a, a)); }
// GLSL normalize(v) = v / length(v)
SOLB_HD float3 normalize(float3 a) {
#if defined(__CUDA_ARCH__)
    const float inv = rsqrtf(dot(a, a));
    return f3(a.x * inv, a.y * inv, a.z * inv);
#else
    float l = length(a); return f3(a.x / l, a.y / l, a.z / l);
#endif
}
// IEEE version (v / sqrt(dot)): primary-ray directions, whose hit ids are compared bit for bit with the reference's
SOLB_HD float3 normalize_ieee(float3 a) { const float l = length(a); return f3(a.x / l, a.y / l, a.z / l); }
SOLB_HD float3 fmin3(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
SOLB_HD float3 fmax3(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
SOLB_HD float3 xyz(float4 a) { return f3(a.x, a.y, a.z); }
SOLB_HD float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
SOLB_HD float signf_glsl(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

// column-major 4x4 (glam::Mat4 / GLSL mat4) times (v, w)
SOLB_HD float3 mat4_mul_point(const float *m, float3 v) {
    return f3(m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12],
              m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13],
              m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14]);
}
SOLB_HD float3 mat4_mul_dir(const float *m, float3 v) {
    return f3(m[0] * v.x + m[4] * v.y + m[8] * v.z,
              m[1] * v.x + m[5] * v.y + m[9] * v.z,
              m[2] * v.x + m[6] * v.y + m[10] * v.z);
}

struct Aabb {
    float3 lo, hi;
};
SOLB_HD float half_area(float3 lo, float3 hi) {
    float3 d = hi - lo;
    return d.x * d.y + d.y * d.z + d.z * d.x;
}

#define SOLB_MISS 0xffffffffu

}  // namespace solb
