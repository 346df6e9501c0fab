// host_math.h — the few glam 0.20.2 (Cargo.lock:509-510) Mat4 operations the hot path needs on the host,
// restated from the crate's scalar-math path (GLM cofactor scheme).  Column-major float[16].
// Call sites replaced: src/ray/mod.rs:29,116 (inverse().transpose()), examples/5-pathtrace.rs:21-28.
#pragma once
#include <math.h>
#include <string.h>

namespace solb {

inline void mat4_identity(float *m) {
    memset(m, 0, 16 * sizeof(float));
    m[0] = m[5] = m[10] = m[15] = 1.0f;
}

// a * b (column-major)
inline void mat4_mul(const float *a, const float *b, float *out) {
    float r[16];
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            float acc = 0.0f;
            for (int k = 0; k < 4; k++) acc = acc + a[4 * k + row] * b[4 * c + k];
            r[4 * c + row] = acc;
        }
    memcpy(out, r, sizeof(r));
}

inline void mat4_transpose(const float *m, float *out) {
    float r[16];
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) r[4 * c + row] = m[4 * row + c];
    memcpy(out, r, sizeof(r));
}

// glam Mat4::inverse
inline void mat4_inverse(const float *m, float *out) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3];
    const float m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    const float m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11];
    const float m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];
    const float coef00 = m22 * m33 - m32 * m23, coef02 = m12 * m33 - m32 * m13, coef03 = m12 * m23 - m22 * m13;
    const float coef04 = m21 * m33 - m31 * m23, coef06 = m11 * m33 - m31 * m13, coef07 = m11 * m23 - m21 * m13;
    const float coef08 = m21 * m32 - m31 * m22, coef10 = m11 * m32 - m31 * m12, coef11 = m11 * m22 - m21 * m12;
    const float coef12 = m20 * m33 - m30 * m23, coef14 = m10 * m33 - m30 * m13, coef15 = m10 * m23 - m20 * m13;
    const float coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12, coef19 = m10 * m22 - m20 * m12;
    const float coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11, coef23 = m10 * m21 - m20 * m11;
    const float fac0[4] = { coef00, coef00, coef02, coef03 }, fac1[4] = { coef04, coef04, coef06, coef07 };
    const float fac2[4] = { coef08, coef08, coef10, coef11 }, fac3[4] = { coef12, coef12, coef14, coef15 };
    const float fac4[4] = { coef16, coef16, coef18, coef19 }, fac5[4] = { coef20, coef20, coef22, coef23 };
    const float vec0[4] = { m10, m00, m00, m00 }, vec1[4] = { m11, m01, m01, m01 };
    const float vec2[4] = { m12, m02, m02, m02 }, vec3[4] = { m13, m03, m03, m03 };
    const float sign_a[4] = { 1, -1, 1, -1 }, sign_b[4] = { -1, 1, -1, 1 };
    float inv[16];
    for (int i = 0; i < 4; i++) {
        inv[0 + i] = (vec1[i] * fac0[i] - vec2[i] * fac1[i] + vec3[i] * fac2[i]) * sign_a[i];
        inv[4 + i] = (vec0[i] * fac0[i] - vec2[i] * fac3[i] + vec3[i] * fac4[i]) * sign_b[i];
        inv[8 + i] = (vec0[i] * fac1[i] - vec1[i] * fac3[i] + vec3[i] * fac5[i]) * sign_a[i];
        inv[12 + i] = (vec0[i] * fac2[i] - vec1[i] * fac4[i] + vec2[i] * fac5[i]) * sign_b[i];
    }
    const float d0 = m00 * inv[0], d1 = m01 * inv[4], d2 = m02 * inv[8], d3 = m03 * inv[12];
    const float dot1 = ((d0 + d1) + d2) + d3;
    const float rcp = 1.0f / dot1;
    for (int i = 0; i < 16; i++) out[i] = inv[i] * rcp;
}

}  // namespace solb
