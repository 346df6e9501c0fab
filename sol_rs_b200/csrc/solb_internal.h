// solb_internal.h — device-side scene/accel views shared by build.cu, trace.cu and solb_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bvh.cuh"

namespace solb {

// One primitive section == one BLAS (src/ray/mod.rs:78-134): geometry ranges in the concatenated arrays.
struct DeviceBlas {
    uint32_t first_vertex;  // section's first vertex in the concatenated vertex array
    uint32_t first_index;   // section's first index in the concatenated index array
    uint32_t n_indices;
    uint32_t first_tri;     // first geometry triangle (= first shading record) of this BLAS
};

// One instance.  The reference creates exactly one per BLAS (src/ray/mod.rs:122 "TODO: support multiple instances
// per BLAS"); solb_scene_add_instance adds further instances of an existing BLAS.
struct DeviceInstance {
    uint32_t first_vertex;  // copy of the BLAS's geometry range
    uint32_t first_index;
    uint32_t n_indices;
    uint32_t material;      // section.material_index (kept for inspection)
    uint32_t blas;          // geometry this instance places
    uint32_t shade_first_tri;  // DeviceBlas.first_tri of that geometry
    float transform[16];    // SceneInstance.transform (column-major)
    float transform_it[16]; // SceneInstance.transform_it = inverse().transpose()
    float mat[12];          // MaterialInfo of materials[gl_InstanceID]: base @0, emissive @4, metallic @8, roughness @9;
                            // @10 (MaterialInfo::padding1) carries bits(texture index + 1), 0 = untextured (solb_scene_set_textures)
};
static_assert(sizeof(DeviceInstance) == 200, "DeviceInstance layout is mirrored by tests/emu_lib.py");

struct DeviceSceneView {
    uint32_t n_instances, n_tris;    // n_tris: triangles over all instances (flattened build)
    const uint32_t *inst_first_tri;  // [n_instances + 1] exclusive prefix of per-instance triangle counts
    const DeviceInstance *instances;
    const float4 *vertices;          // reference ModelVertex array, 4 x float4 per vertex: pos, color, normal, uv
    const uint32_t *indices;         // section-relative u32 indices
    uint32_t n_blas, n_geom_tris;    // unique geometry: BLAS count, triangles over all BLASes
    const DeviceBlas *blas;          // [n_blas], first_tri ascending
};

// 112-byte shading record per global triangle: the three ModelVertex {pos, normal, color} triples the
// closest-hit shaders fetch through indices[]/vertices[] (pathtrace.rchit:61-67), de-indexed once at
// scene creation so a hit costs 7 contiguous 16-byte loads instead of 3 index + 3x64 B scattered ones.
//   f[9*k + 0..2] = pos_k, f[9*k + 3..5] = normal_k, f[9*k + 6..8] = color_k.rgb, f[27] = 0
struct ShadeRecord {
    float4 q[7];
};
static_assert(sizeof(ShadeRecord) == 112, "ShadeRecord must be 112 bytes");

// Base-colour textures.  The reference reserves SceneInstance::texture_offset (src/ray/mod.rs:20,
// pathtrace.rchit:31) and never reads it; solb_scene_set_textures gives it a meaning: the instance's albedo is multiplied by
// a bilinear sample of texture `texture_offset` at the interpolated ModelVertex::uv.  Texels are linear-light float4 (sRGB
// decoded on upload), rows top first (glTF: uv (0, 0) = top-left texel corner).  No textures bound (n_tex = 0): the
// reference's shading, bit for bit.
struct TexDesc {
    const float4 *texels;
    uint32_t width, height;
    uint32_t wrap_s, wrap_t;  // glTF sampler codes: 10497 repeat, 33071 clamp to edge, 33648 mirrored repeat
};
static_assert(sizeof(TexDesc) == 24, "TexDesc");
struct TexBinding {
    const TexDesc *tex = nullptr;
    const float4 *vertices = nullptr;    // ModelVertex array (4 x float4: pos, color, normal, uv), as the closest-hit shader reads it
    const uint32_t *indices = nullptr;   // section-relative indices
    uint32_t n_tex = 0, pad = 0;
};

// Per-launch constants: the parts of SceneUniforms the RT stages read (view_inverse,
// projection_inverse, frame.z: pathtrace.rgen:13-21) + push constant + specialization constant +
// the shader literals exposed as parameters (SolbTraceParams).
struct FrameConsts {
    float view_inv[16];
    float proj_inv[16];
    float3 origin;        // view_inverse * (0,0,0,1), identical for every pixel (pathtrace.rgen:55)
    float tmin, tmax;     // preparePayload: max(1, |origin|) * 1e-3, 1e4 (pathtrace.rgen:35)
    uint32_t width, height, frame;
    // image rows this launch covers (tile split): n_bands bands of band_rows rows, band k starting at row
    // row_begin + k * band_stride, clipped to the image.  Whole image: one band of `height` rows at 0.
    uint32_t row_begin, band_rows, band_stride, n_bands;
    int32_t accum_start;
    uint32_t enable_sky, spp, max_bounces, accum_mode;
    TexBinding texb;      // base-colour textures of the scene (n_tex = 0: none)
};

struct AccelStorage {
    uint4 *nodes_u4() const { return (uint4 *)nodes; }
    float4 *tris_f4() const { return (float4 *)tris; }
    float4 *inst_leaves_f4() const { return (float4 *)inst_leaves; }
    Node8 *nodes = nullptr;
    Tri48 *tris = nullptr;
    size_t nodes_cap = 0, tris_cap = 0;  // elements allocated (flattened builds: a rebuild reuses arrays that are large enough)
    uint32_t n_wide = 0, n_tris = 0, depth = 0, n_binary = 0;
    float sah_lbvh = 0.0f, sah_final = 0.0f;
    float lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
    // two-level mode (SOLB_ACCEL_TWO_LEVEL): nodes[0, tlas_cap) = TLAS region (root = node 0, rebuilt in place by
    // rebuild_tlas), nodes[tlas_cap + b] = root of BLAS b, then the BLAS interiors; tris are object-space.
    bool two_level = false;
    InstLeaf *inst_leaves = nullptr;   // [n_instances] TLAS leaf records in TLAS leaf order
    float4 *blas_box = nullptr;        // [2 * n_blas] object-space root boxes (lo, hi)
    uint32_t tlas_cap = 0, n_blas = 0, n_tlas_wide = 0, tlas_depth = 0, blas_depth = 0;
    void *d_tlas_info = nullptr, *h_tlas_info = nullptr;  // single-CTA TLAS build: result block (device) + pinned mirror
    void *d_tlas_dp = nullptr;                              // ... and its optimal-collapse table (32 B per binary node)
    void release() {
        if (d_tlas_info) cudaFree(d_tlas_info);
        if (h_tlas_info) cudaFreeHost(h_tlas_info);
        if (d_tlas_dp) cudaFree(d_tlas_dp);
        d_tlas_info = h_tlas_info = d_tlas_dp = nullptr;
        if (nodes) cudaFree(nodes);
        if (tris) cudaFree(tris);
        if (inst_leaves) cudaFree(inst_leaves);
        if (blas_box) cudaFree(blas_box);
        nodes = nullptr; tris = nullptr; inst_leaves = nullptr; blas_box = nullptr;
        nodes_cap = tris_cap = 0;
        n_wide = n_tris = depth = n_binary = 0;
        two_level = false;
        tlas_cap = n_blas = n_tlas_wide = tlas_depth = blas_depth = 0;
    }
};

struct BuildOptions {
    int treelet_passes = 2;
    int treelet_gamma = 7;
    int coop_treelet = 1;  // warp-cooperative treelet kernel (0: per-thread reference version)
    int dp_collapse = 1;   // SAH-optimal wide collapse (Ylitie et al. 2017); 0: greedy largest-area-first
    float dp_cn = 2.0f, dp_cp = 1.0f;  // its cost model: 8-wide node visit, primitive test (per unit of box area)
    int dp_max_leaf = SOLB_MAX_LEAF_TRIS;  // ... and the largest leaf it may form
    int ploc = 0;          // flattened build: PLOC over the Morton-sorted leaves instead of LBVH + treelet restructuring
};

cudaError_t build_accel(cudaStream_t st, const DeviceSceneView &sv, AccelStorage &out, const BuildOptions &opt, uint64_t *launches);
// Two-level build: every BLAS in object space in ONE batched pass (segmented Morton keys), then the TLAS over the
// instances' world boxes.  rebuild_tlas redoes only the second half (TLAS::regenerate, src/ray/acceleration.rs:402-467).
cudaError_t build_accel_two_level(cudaStream_t st, const DeviceSceneView &sv, AccelStorage &out, const BuildOptions &opt,
                                  uint64_t *launches);
cudaError_t rebuild_tlas(cudaStream_t st, const DeviceSceneView &sv, AccelStorage &out, uint64_t *launches);
// stream-ordered scratch of the calling thread's next builds comes from `pool` (nullptr: the device's default pool)
void set_build_pool(cudaMemPool_t pool);
cudaError_t sort_pairs_device(cudaStream_t st, uint64_t *keys, uint32_t *vals, uint32_t n, int key_bits);

}  // namespace solb
