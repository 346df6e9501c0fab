//! Raw bindings to libsolb.so — one declaration per SOLB_API entry point of include/solb.h, same order.
//! tests/test_host_and_abi.py diffs this block against the header symbol for symbol (the build image has no rustc).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct solb_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct solb_scene {
    _private: [u8; 0],
}
#[repr(C)]
pub struct solb_target {
    _private: [u8; 0],
}
#[repr(C)]
pub struct solb_fence {
    _private: [u8; 0],
}

/// identical bytes to sol::scene::ModelVertex (src/scene/mesh.rs:9-14), 64 B
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SolbModelVertex {
    pub pos: [f32; 4],
    pub color: [f32; 4],
    pub normal: [f32; 4],
    pub uv: [f32; 4],
}
/// identical bytes to sol::scene::MaterialInfo (src/scene/mod.rs:19-29), 48 B
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SolbMaterialInfo {
    pub base_color: [f32; 4],
    pub emissive: [f32; 3],
    pub padding0: f32,
    pub metallic: f32,
    pub roughness: f32,
    pub padding1: f32,
    pub padding2: f32,
}
/// identical bytes to sol::ray::SceneInstance (src/ray/mod.rs:16-24), 144 B
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct SolbSceneInstance {
    pub id: u32,
    pub texture_offset: u32,
    pub padding: [f32; 2],
    pub transform: [f32; 16],
    pub transform_it: [f32; 16],
}
/// identical bytes to SceneUniforms (examples/5-pathtrace.rs:7-17), padded to 400 B
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct SolbSceneUniforms {
    pub model: [f32; 16],
    pub view: [f32; 16],
    pub view_inverse: [f32; 16],
    pub projection: [f32; 16],
    pub projection_inverse: [f32; 16],
    pub model_view_projection: [f32; 16],
    pub frame: [u32; 3],
    pub _pad: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SolbSection {
    pub first_vertex: u32,
    pub n_vertices: u32,
    pub first_index: u32,
    pub n_indices: u32,
    pub material_index: u32,
}
pub const SOLB_NO_TEXTURE: u32 = 0xffff_ffff;
/// base-colour texture handed to solb_scene_set_textures: rgba8 rows top first
#[repr(C)]
pub struct SolbTextureDesc {
    pub rgba8: *const u8,
    pub width: u32,
    pub height: u32,
    pub wrap_s: u32,
    pub wrap_t: u32,
    pub srgb: u32,
    pub _pad: u32,
}
#[repr(C)]
pub struct SolbMeshDesc {
    pub vertices: *const SolbModelVertex,
    pub n_vertices: u32,
    pub indices: *const u32,
    pub n_indices: u32,
    pub sections: *const SolbSection,
    pub n_sections: u32,
    pub transform: [f32; 16],
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SolbTraceParams {
    pub accum_start_frame: i32,
    pub enable_sky: u32,
    pub samples_per_frame: u32,
    pub max_bounces: u32,
    pub schedule: u32,
    pub accum_mode: u32,
    pub collect_stats: u32,
    pub tile_row_begin: u32,
    pub tile_row_count: u32,
    pub tile_row_stride: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SolbStats {
    pub rays: u64,
    pub hits: u64,
    pub paths: u64,
    pub nodes_visited: u64,
    pub tris_tested: u64,
    pub kernel_launches: u64,
    pub last_build_ms: f32,
    pub last_trace_ms: f32,
    pub trace_kernel_ms_total: f32,
    pub trace_kernel_launches: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SolbAccelInfo {
    pub n_instances: u32,
    pub n_triangles: u32,
    pub n_wide_nodes: u32,
    pub wide_depth: u32,
    pub n_binary_nodes: u32,
    pub sah_cost_binary: f32,
    pub sah_cost_lbvh: f32,
    pub scene_lo: [f32; 3],
    pub scene_hi: [f32; 3],
    pub mode: u32,
    pub n_blas: u32,
    pub n_tlas_nodes: u32,
    pub tlas_depth: u32,
}

pub const SOLB_OK: c_int = 0;
pub const SOLB_FORMAT_RGBA32F: u32 = 0;
pub const SOLB_FORMAT_RGBA8: u32 = 1;
pub const SOLB_FORMAT_RG32UI: u32 = 2;
pub const SOLB_SCHEDULE_WAVEFRONT: u32 = 0;
pub const SOLB_SCHEDULE_MEGAKERNEL: u32 = 1;
pub const SOLB_SCHEDULE_AUTO: u32 = 2;
pub const SOLB_SCHEDULE_WARPFRONT: u32 = 3;
pub const SOLB_ACCEL_FLAT: u32 = 0;
pub const SOLB_ACCEL_TWO_LEVEL: u32 = 1;
pub const SOLB_ACCUM_MIX: u32 = 0;
pub const SOLB_ACCUM_SUM: u32 = 1;
pub const SOLB_COMM_ID_BYTES: usize = 128;

extern "C" {
    // ---- context ----
    pub fn solb_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut solb_ctx) -> c_int;
    pub fn solb_ctx_preload(ctx: *mut solb_ctx) -> c_int;
    pub fn solb_ctx_destroy(ctx: *mut solb_ctx) -> c_int;
    pub fn solb_synchronize(ctx: *mut solb_ctx) -> c_int;
    pub fn solb_fence_create(ctx: *mut solb_ctx, out: *mut *mut solb_fence) -> c_int;
    pub fn solb_fence_signal(fence: *mut solb_fence) -> c_int;
    pub fn solb_fence_wait(fence: *mut solb_fence) -> c_int;
    pub fn solb_fence_destroy(fence: *mut solb_fence) -> c_int;
    pub fn solb_host_alloc(ctx: *mut solb_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn solb_host_free(ctx: *mut solb_ctx, ptr: *mut c_void) -> c_int;
    pub fn solb_ctx_trim(ctx: *mut solb_ctx) -> c_int;
    pub fn solb_last_error(ctx: *mut solb_ctx) -> *const c_char;
    pub fn solb_version() -> u32;
    pub fn solb_stats_get(ctx: *mut solb_ctx, out: *mut SolbStats) -> c_int;
    pub fn solb_stats_reset(ctx: *mut solb_ctx) -> c_int;
    pub fn solb_set_timing(ctx: *mut solb_ctx, enabled: c_int) -> c_int;
    // ---- scene / acceleration structure ----
    pub fn solb_scene_create(ctx: *mut solb_ctx, meshes: *const SolbMeshDesc, n_meshes: u32, materials: *const SolbMaterialInfo,
                             n_materials: u32, out: *mut *mut solb_scene) -> c_int;
    pub fn solb_scene_destroy(scene: *mut solb_scene) -> c_int;
    pub fn solb_accel_build(scene: *mut solb_scene) -> c_int;
    pub fn solb_instance_set_transform(scene: *mut solb_scene, index: u32, transform: *const f32) -> c_int;
    pub fn solb_scene_update(scene: *mut solb_scene) -> c_int;
    pub fn solb_tlas_regenerate(scene: *mut solb_scene) -> c_int;
    pub fn solb_scene_set_textures(scene: *mut solb_scene, textures: *const SolbTextureDesc, n_textures: u32, material_texture: *const u32,
                                   n_materials: u32) -> c_int;
    pub fn solb_scene_add_instance(scene: *mut solb_scene, source_instance: u32, transform: *const f32, material_index: u32,
                                   out_index: *mut u32) -> c_int;
    pub fn solb_scene_set_accel_mode(scene: *mut solb_scene, mode: u32) -> c_int;
    pub fn solb_scene_instance_count(scene: *mut solb_scene, out: *mut u32) -> c_int;
    pub fn solb_scene_get_instances(scene: *mut solb_scene, out: *mut SolbSceneInstance, capacity: u32) -> c_int;
    pub fn solb_scene_instance_triangles(scene: *mut solb_scene, out: *mut u32, capacity: u32) -> c_int;
    pub fn solb_accel_info(scene: *mut solb_scene, out: *mut SolbAccelInfo) -> c_int;
    pub fn solb_accel_read_nodes(scene: *mut solb_scene, host: *mut c_void, bytes: usize) -> c_int;
    pub fn solb_accel_read_triangles(scene: *mut solb_scene, host: *mut c_void, bytes: usize) -> c_int;
    // ---- targets ----
    pub fn solb_target_create(ctx: *mut solb_ctx, width: u32, height: u32, format: u32, out: *mut *mut solb_target) -> c_int;
    pub fn solb_target_destroy(t: *mut solb_target) -> c_int;
    pub fn solb_target_clear(t: *mut solb_target) -> c_int;
    pub fn solb_target_readback(t: *mut solb_target, host: *mut c_void, bytes: usize) -> c_int;
    pub fn solb_target_readback_async(t: *mut solb_target, host: *mut c_void, bytes: usize) -> c_int;
    pub fn solb_target_upload(t: *mut solb_target, host: *const c_void, bytes: usize) -> c_int;
    pub fn solb_target_device_ptr(t: *mut solb_target, out: *mut *mut c_void) -> c_int;
    pub fn solb_target_info(t: *mut solb_target, width: *mut u32, height: *mut u32, format: *mut u32) -> c_int;
    // ---- trace launches ----
    pub fn solb_trace_params_default(p: *mut SolbTraceParams, pipeline: c_int);
    pub fn solb_trace_pathtrace(scene: *mut solb_scene, uniforms: *const SolbSceneUniforms, params: *const SolbTraceParams,
                                accum: *mut solb_target, render: *mut solb_target) -> c_int;
    pub fn solb_set_blue_noise(ctx: *mut solb_ctx, rgba8: *const u8, width: u32, height: u32) -> c_int;
    pub fn solb_trace_ao(scene: *mut solb_scene, uniforms: *const SolbSceneUniforms, params: *const SolbTraceParams,
                         image: *mut solb_target) -> c_int;
    pub fn solb_trace_debug(scene: *mut solb_scene, uniforms: *const SolbSceneUniforms, render: *mut solb_target,
                            ids: *mut solb_target, attribs: *mut solb_target) -> c_int;
    pub fn solb_trace_rays(scene: *mut solb_scene, rays: *const f32, n: u32, hits: *mut u32, t_out: *mut f32) -> c_int;
    pub fn solb_test_sort_pairs(ctx: *mut solb_ctx, keys: *mut u64, values: *mut u32, n: u32, key_bits: c_int) -> c_int;
    // ---- multi-GPU ----
    pub fn solb_resolve_sum(ctx: *mut solb_ctx, sum: *mut solb_target, accum_out: *mut solb_target, render: *mut solb_target) -> c_int;
    pub fn solb_comm_unique_id(id_out: *mut u8) -> c_int;
    pub fn solb_comm_init(ctx: *mut solb_ctx, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn solb_comm_info(ctx: *mut solb_ctx, rank: *mut c_int, world: *mut c_int, nccl_version: *mut c_int) -> c_int;
    pub fn solb_comm_destroy(ctx: *mut solb_ctx) -> c_int;
    pub fn solb_reduce_accum(ctx: *mut solb_ctx, sum: *mut solb_target, root: c_int, accum_out: *mut solb_target,
                             render: *mut solb_target) -> c_int;
    pub fn solb_allgather_rows(ctx: *mut solb_ctx, target: *mut solb_target, band_rows: u32) -> c_int;
}
