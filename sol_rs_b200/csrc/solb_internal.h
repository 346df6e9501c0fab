// solb_internal.h — device-side scene/accel views shared by build.cu, trace.cu and solb_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bvh.cuh"

namespace solb {

// One reference instance == one primitive section == one BLAS (src/ray/mod.rs:78-134).
struct DeviceInstance {
    uint32_t first_vertex;  // section's first vertex in the concatenated vertex array
    uint32_t first_index;   // section's first index in the concatenated index array
    uint32_t n_indices;
    uint32_t material;      // section.material_index (kept for inspection)
    float transform[16];    // SceneInstance.transform (column-major)
    float transform_it[16]; // SceneInstance.transform_it = inverse().transpose()
    float mat[12];          // MaterialInfo of materials[gl_InstanceID]: base @0, emissive @4, metallic @8, roughness @9
};

struct DeviceSceneView {
    uint32_t n_instances, n_tris;
    const uint32_t *inst_first_tri;  // [n_instances + 1] exclusive prefix of triangle counts
    const DeviceInstance *instances;
    const float4 *vertices;          // reference ModelVertex array, 4 x float4 per vertex: pos, color, normal, uv
    const uint32_t *indices;         // section-relative u32 indices
};

// 112-byte shading record per global triangle: the three ModelVertex {pos, normal, color} triples the
// closest-hit shaders fetch through indices[]/vertices[] (pathtrace.rchit:61-67), de-indexed once at
// scene creation so a hit costs 7 contiguous 16-byte loads instead of 3 index + 3x64 B scattered ones.
//   f[9*k + 0..2] = pos_k, f[9*k + 3..5] = normal_k, f[9*k + 6..8] = color_k.rgb, f[27] = 0
struct ShadeRecord {
    float4 q[7];
};
static_assert(sizeof(ShadeRecord) == 112, "ShadeRecord must be 112 bytes");

// Per-launch constants: the parts of SceneUniforms the RT stages read (view_inverse,
// projection_inverse, frame.z: pathtrace.rgen:13-21) + push constant + specialization constant +
// the shader literals exposed as parameters (SolbTraceParams).
struct FrameConsts {
    float view_inv[16];
    float proj_inv[16];
    float3 origin;        // view_inverse * (0,0,0,1), identical for every pixel (pathtrace.rgen:55)
    float tmin, tmax;     // preparePayload: max(1, |origin|) * 1e-3, 1e4 (pathtrace.rgen:35)
    uint32_t width, height, frame;
    int32_t accum_start;
    uint32_t enable_sky, spp, max_bounces, accum_mode;
};

struct AccelStorage {
    uint4 *nodes_u4() const { return (uint4 *)nodes; }
    float4 *tris_f4() const { return (float4 *)tris; }
    Node8 *nodes = nullptr;
    Tri48 *tris = nullptr;
    uint32_t n_wide = 0, n_tris = 0, depth = 0, n_binary = 0;
    float sah_lbvh = 0.0f, sah_final = 0.0f;
    float lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
    void release() {
        if (nodes) cudaFree(nodes);
        if (tris) cudaFree(tris);
        nodes = nullptr; tris = nullptr;
        n_wide = n_tris = depth = n_binary = 0;
    }
};

struct BuildOptions {
    int treelet_passes = 2;
    int treelet_gamma = 7;
    int coop_treelet = 1;  // warp-cooperative treelet kernel (0: per-thread reference version)
};

cudaError_t build_accel(cudaStream_t st, const DeviceSceneView &sv, AccelStorage &out, const BuildOptions &opt, uint64_t *launches);
cudaError_t sort_pairs_device(cudaStream_t st, uint64_t *keys, uint32_t *vals, uint32_t n, int key_bits);

}  // namespace solb
