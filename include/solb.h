/*
 * solb.h — C ABI of libsolb.so, the B200 (sm_100a) replacement for the ray-tracing hot path of
 * num3ric/sol-rs.  Everything below is `extern "C"`, POD-only, plain pointers and sizes.
 *
 * The reference has no FFI/plugin boundary of its own (it is a Rust crate that drives the Vulkan
 * driver directly), so each entry point names the reference call(s) a Rust shim
 * (rust/sol/) replaces with it.  Paths are relative to the reference root.
 *
 * Conventions
 *   - every function returns int: SOLB_OK (0) or a negative SolbStatus; text via solb_last_error().
 *   - never aborts / throws across the ABI.  The Rust shim turns non-zero into panic!() to keep
 *     the reference's unwrap()/expect() behaviour (src/scene/mod.rs:140, src/ray/pipeline.rs:105).
 *   - one ctx = one CUDA device + one stream; a ctx is not thread-safe (the reference is
 *     single-threaded: src/lib.rs:174).  Multi-GPU = one ctx per device / process.
 *   - everything is ORDERED as if enqueued on that one stream.  Internally the path tracer's frame kernels run on side
 *     streams of the ctx so that consecutive frames overlap (frames in flight, like the reference's per-swapchain-image
 *     command buffers); their resolves into the targets, and every other call, stay in order on the ctx stream, so work
 *     the caller enqueues on that stream after a call sees its result.
 *   - host pointers passed in are copied during the call and never retained
 *     (reference: Buffer::from_data copies, src/buffer.rs:186-206).
 *   - matrices are column-major float[16] exactly as glam::Mat4 lays them out.
 *   - there is NO CPU fallback: without a CUDA device every entry point fails with SOLB_ERR_CUDA.
 */
#ifndef SOLB_H
#define SOLB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SOLB_API __declspec(dllexport)
#else
#define SOLB_API __attribute__((visibility("default")))
#endif

#define SOLB_VERSION 0x000100

typedef enum SolbStatus {
    SOLB_OK = 0,
    SOLB_ERR_INVALID = -1,      /* bad argument / handle / size                         */
    SOLB_ERR_CUDA = -2,         /* CUDA runtime error (no device, launch failure, OOM)   */
    SOLB_ERR_NOT_BUILT = -3,    /* trace before solb_accel_build                         */
    SOLB_ERR_UNSUPPORTED = -4,  /* e.g. non-indexed primitive (SURVEY App.A item 4)      */
    SOLB_ERR_OVERFLOW = -5      /* internal capacity (traversal stack depth, queue)      */
} SolbStatus;

typedef struct solb_ctx solb_ctx;       /* replaces sol::Context for this path: src/context.rs        */
typedef struct solb_scene solb_scene;   /* replaces ray::SceneDescription (+BLAS/TLAS): src/ray/mod.rs */
typedef struct solb_target solb_target; /* replaces sol::Image2d storage targets: src/texture.rs:36-96 */
typedef struct solb_fence solb_fence;   /* replaces AppFrameData::in_flight_fence: src/renderer.rs:8      */

/* ---- POD data contracts (identical bytes to the reference) ------------- */

/* src/scene/mesh.rs:9-14; assets/glsl/pathtrace.rchit:11-16 */
typedef struct SolbModelVertex { float pos[4], color[4], normal[4], uv[4]; } SolbModelVertex; /* 64 B */

/* src/scene/mod.rs:19-29; assets/glsl/pathtrace.rchit:18-26 */
typedef struct SolbMaterialInfo {
    float base_color[4];
    float emissive[3];
    float padding0;
    float metallic, roughness, padding1, padding2;
} SolbMaterialInfo; /* 48 B */

/* src/ray/mod.rs:16-24; assets/glsl/pathtrace.rchit:28-35 */
typedef struct SolbSceneInstance {
    uint32_t id, texture_offset;
    float padding[2];
    float transform[16];
    float transform_it[16];
} SolbSceneInstance; /* 144 B */

/* examples/5-pathtrace.rs:7-17; assets/glsl/pathtrace.rgen:13-21.  Only view_inverse,
 * projection_inverse and frame[2] are read by the ray-tracing stages. */
typedef struct SolbSceneUniforms {
    float model[16], view[16], view_inverse[16], projection[16], projection_inverse[16], model_view_projection[16];
    uint32_t frame[3]; /* (width, height, elapsed_ticks) */
    uint32_t _pad;
} SolbSceneUniforms; /* 400 B */

/* scene::PrimitiveSection (src/scene/mod.rs:37-44): one section -> one BLAS -> one instance */
typedef struct SolbSection {
    uint32_t first_vertex, n_vertices; /* BufferPart vertices (ModelVertex units, within the mesh)  */
    uint32_t first_index, n_indices;   /* BufferPart indices (u32 units, within the mesh)           */
    uint32_t material_index;           /* into the scene's material array                           */
} SolbSection;

/* scene::Mesh (src/scene/mesh.rs:53-61) as host arrays */
typedef struct SolbMeshDesc {
    const SolbModelVertex *vertices; uint32_t n_vertices;
    const uint32_t *indices; uint32_t n_indices; /* section-relative, as glTF stores them */
    const SolbSection *sections; uint32_t n_sections;
    float transform[16]; /* Mesh::transform, column-major */
} SolbMeshDesc;

typedef enum SolbTargetFormat {
    SOLB_FORMAT_RGBA32F = 0, /* vk R32G32B32A32_SFLOAT: accumulation image, examples/5-pathtrace.rs:222 */
    SOLB_FORMAT_RGBA8 = 1,   /* vk R8G8B8A8_UNORM: render image, examples/5-pathtrace.rs:229           */
    SOLB_FORMAT_RG32UI = 2   /* (instance, primitive) ids of the primary hit; 0xffffffff = miss        */
} SolbTargetFormat;

typedef enum SolbSchedule {
    SOLB_SCHEDULE_WAVEFRONT = 0, /* queue-based wavefront: persistent trace kernel + shade/compact kernels */
    SOLB_SCHEDULE_MEGAKERNEL = 1,/* one thread per pixel, flattened bounce loop                           */
    SOLB_SCHEDULE_AUTO = 2,      /* megakernel for tiny hierarchies (<= 8 wide nodes: no traversal divergence to
                                    hide, queue traffic dominates), warp-local wavefront otherwise           */
    SOLB_SCHEDULE_WARPFRONT = 3  /* warp-local wavefront: ONE persistent kernel per frame; every warp traces its own pool of
                                    pixels with the voted traversal and shades its finished rays 32 wide (no global queues,
                                    no waves, no host polling)                                               */
} SolbSchedule;

typedef enum SolbAccelMode {
    SOLB_ACCEL_FLAT = 0,     /* default: instance transforms baked into world-space triangles, ONE hierarchy (the reference has
                                exactly one instance per BLAS, src/ray/mod.rs:122, so instancing saves nothing there)       */
    SOLB_ACCEL_TWO_LEVEL = 1 /* TLAS of instances over shared object-space BLASes (src/ray/acceleration.rs:136-239,344-400):
                                instances of one BLAS share its nodes / triangles, solb_tlas_regenerate rebuilds the TLAS only */
} SolbAccelMode;

typedef enum SolbAccumMode {
    SOLB_ACCUM_MIX = 0, /* reference: new = mix(old, frame, 1/(frame+1-start)), pathtrace.rgen:89-101 */
    SOLB_ACCUM_SUM = 1  /* multi-GPU: accum.xyz += frame colour, accum.w += 1 (resolved later)         */
} SolbAccumMode;

/* Everything the reference passes besides the uniform block: the push constant
 * (examples/5-pathtrace.rs:306-314), specialization constant 0 (:103) and the shader literals
 * (pathtrace.rgen:43-44; ao.rgen:42-43), which BASELINE.json's configs need as parameters.
 * solb_trace_params_default() fills the reference values. */
typedef struct SolbTraceParams {
    int32_t accum_start_frame;  /* push.accum_start_frame                              */
    uint32_t enable_sky;        /* ENABLE_SKYLIGHT specialization constant             */
    uint32_t samples_per_frame; /* sampleCount literal: 8 (pathtrace), 4 (ao)          */
    uint32_t max_bounces;       /* maxBounces literal: 32 (pathtrace); ao: max_samples 4 */
    uint32_t schedule;          /* SolbSchedule                                        */
    uint32_t accum_mode;        /* SolbAccumMode                                       */
    uint32_t collect_stats;     /* 1: instrumented kernels count nodes / triangles per ray */
    uint32_t tile_row_begin;    /* tile split of ONE frame across GPUs: trace only image rows               */
    uint32_t tile_row_count;    /* [tile_row_begin, tile_row_begin + tile_row_count) of the full-size targets; 0 = all.
                                   Seeds, pixel coordinates and addresses stay those of the full image, so the union of
                                   the tiles is bit-identical to the undivided frame.                                     */
    uint32_t tile_row_stride;   /* 0: that one band.  Otherwise the band repeats every tile_row_stride rows down to the
                                   bottom of the image (interleaved tiles balance uneven images across ranks).            */
} SolbTraceParams;

typedef struct SolbStats {
    uint64_t rays;            /* traceRayEXT-equivalents since last reset                     */
    uint64_t hits;            /* closest-hit invocations                                      */
    uint64_t paths;           /* samples started                                              */
    uint64_t nodes_visited;   /* 8-wide nodes fetched   (collect_stats only)                  */
    uint64_t tris_tested;     /* triangle records fetched (collect_stats only)                */
    uint64_t kernel_launches; /* kernels launched by this ctx                                 */
    float last_build_ms;      /* solb_accel_build / tlas_regenerate device time               */
    float last_trace_ms;      /* last solb_trace_* device time (only when timing is enabled)  */
    /* timing mode only: cudaEvent pairs around every launch of the dominant (traversal) kernel */
    float trace_kernel_ms_total;     /* summed duration of k_wf_trace / k_pathtrace_mega launches since reset */
    uint32_t trace_kernel_launches;  /* number of those launches                                              */
} SolbStats;

typedef struct SolbAccelInfo {
    uint32_t n_instances, n_triangles;
    uint32_t n_wide_nodes;    /* 80-byte 8-wide nodes               */
    uint32_t wide_depth;      /* depth of the 8-wide tree           */
    uint32_t n_binary_nodes;
    float sah_cost_binary;    /* SAH cost of the binary tree (after treelet pass)      */
    float sah_cost_lbvh;      /* SAH cost of the raw LBVH (before treelet pass)        */
    float scene_lo[3], scene_hi[3];
    uint32_t mode;            /* SolbAccelMode the structure was built with                */
    uint32_t n_blas;          /* unique geometries (primitive sections)                    */
    uint32_t n_tlas_nodes;    /* two-level: 8-wide nodes of the TLAS (0 when flattened)    */
    uint32_t tlas_depth;      /* two-level: depth of the TLAS part of wide_depth           */
} SolbAccelInfo;

/* ---- context ------------------------------------------------------------------------------ */

/* Replaces Context/SharedContext creation (src/context.rs:239-369) for this path.
 * `stream` is a cudaStream_t to launch on (e.g. torch's current stream), or NULL to create a private non-blocking one.
 * NULL never means "the default stream": to adopt the legacy default stream (handle 0 in the runtime API) pass
 * cudaStreamLegacy ((cudaStream_t)0x1), to adopt the per-thread default stream cudaStreamPerThread ((cudaStream_t)0x2). */
SOLB_API int solb_ctx_create(int device, void *stream, solb_ctx **out);
/* Counterpart of the shader compilation in Pipeline::new (src/ray/pipeline.rs:61-115): loads every kernel of the library and
 * creates the scratch pool by pushing a two-triangle scene through all launch paths once, so the first solb_accel_build /
 * solb_trace_* of the process does not pay for CUDA's lazy module loading.  Optional; idempotent. */
SOLB_API int solb_ctx_preload(solb_ctx *ctx);
SOLB_API int solb_ctx_destroy(solb_ctx *ctx);
SOLB_API int solb_synchronize(solb_ctx *ctx); /* queue_wait_idle, src/context.rs:539-559 */
/* Frames in flight.  Replaces the per-frame fence of the reference's render loop: vkCreateFence(SIGNALED) per swapchain image
 * (src/renderer.rs:72-81), queue_submit(..., in_flight_fence) (src/renderer.rs:310-317) and wait_for_and_reset_fence at the
 * start of the frame that reuses the slot (src/renderer.rs:123-131, 188).  A new fence is signalled; solb_fence_signal
 * re-arms it to complete when everything enqueued on the context so far (traces, resolves, asynchronous readbacks) has
 * finished; solb_fence_wait blocks the calling thread until then. */
SOLB_API int solb_fence_create(solb_ctx *ctx, solb_fence **out);
SOLB_API int solb_fence_signal(solb_fence *fence);
SOLB_API int solb_fence_wait(solb_fence *fence);
SOLB_API int solb_fence_destroy(solb_fence *fence);
/* Page-locked host memory for uploads / read-backs (the reference's CpuToGpu / GpuToCpu staging buffers,
 * src/buffer.rs:76-83). */
SOLB_API int solb_host_alloc(solb_ctx *ctx, size_t bytes, void **out);
SOLB_API int solb_host_free(solb_ctx *ctx, void *ptr);
/* Return cached build scratch to the driver (the stream-ordered pool keeps it between builds so per-frame rebuilds do not
 * re-map memory; cf. the reference's gpu-allocator blocks, src/context.rs:229). Synchronises. */
SOLB_API int solb_ctx_trim(solb_ctx *ctx);
SOLB_API const char *solb_last_error(solb_ctx *ctx); /* ctx may be NULL: last global error */
SOLB_API uint32_t solb_version(void);
SOLB_API int solb_stats_get(solb_ctx *ctx, SolbStats *out); /* synchronises */
SOLB_API int solb_stats_reset(solb_ctx *ctx);
SOLB_API int solb_set_timing(solb_ctx *ctx, int enabled); /* cudaEvent pair around each trace/build (renderer.rs:204-225 idea) */

/* ---- scene / acceleration structure -------------------------------------------------------- */

/* Replaces SceneDescription::from_meshes up to (not including) the BLAS/TLAS builds
 * (src/ray/mod.rs:59-156): flattens meshes x sections into instances (id = running count),
 * uploads vertices / indices / per-instance material + transform. */
SOLB_API int solb_scene_create(solb_ctx *ctx, const SolbMeshDesc *meshes, uint32_t n_meshes,
                               const SolbMaterialInfo *materials, uint32_t n_materials, solb_scene **out);
SOLB_API int solb_scene_destroy(solb_scene *scene);

/* Base-colour textures (an extension: the reference reserves SceneInstance::texture_offset, src/ray/mod.rs:20 and
 * assets/glsl/pathtrace.rchit:31, and samples nothing).  textures[t].rgba8 = width x height x 4 bytes, rows top first (glTF
 * image order: uv (0, 0) is the top-left corner), copied during the call; srgb != 0 decodes the colour channels with the sRGB
 * transfer function (what glTF prescribes for baseColorTexture).  material_texture[m] = texture of material m or
 * SOLB_NO_TEXTURE.  Afterwards every instance's texture_offset is the texture of its material (solb_scene_get_instances shows
 * it) and 5-pathtrace multiplies the albedo (base_color x vertex colour, pathtrace.rchit:99,109) by a bilinear sample at the
 * interpolated ModelVertex::uv.  n_textures = 0 unbinds: the reference's shading, bit for bit.  Synchronises. */
#define SOLB_NO_TEXTURE 0xffffffffu
typedef struct SolbTextureDesc {
    const uint8_t *rgba8;
    uint32_t width, height;
    uint32_t wrap_s, wrap_t; /* glTF sampler codes: 10497 REPEAT (also 0), 33071 CLAMP_TO_EDGE, 33648 MIRRORED_REPEAT */
    uint32_t srgb;
    uint32_t _pad;
} SolbTextureDesc;
SOLB_API int solb_scene_set_textures(solb_scene *scene, const SolbTextureDesc *textures, uint32_t n_textures,
                                     const uint32_t *material_texture, uint32_t n_materials);

/* Replaces BLAS::new per section + TLAS::new (src/ray/acceleration.rs:136-239,344-400): GPU build of
 * the acceleration structure (Morton LBVH -> treelet SAH -> 8-wide compressed nodes). */
SOLB_API int solb_accel_build(solb_scene *scene);

/* Replaces SceneDescription::blas_transform (src/ray/mod.rs:162-167): sets instance `index`'s
 * transform (and transform_it = inverse().transpose()) on the host copy and marks the TLAS dirty. */
SOLB_API int solb_instance_set_transform(solb_scene *scene, uint32_t index, const float transform[16]);
/* Replaces SceneDescription::update (src/ray/mod.rs:190-192): uploads the instance buffer. */
SOLB_API int solb_scene_update(solb_scene *scene);
/* Replaces SceneDescription::tlas_regenerate -> TLAS::regenerate (src/ray/acceleration.rs:402-467).
 * The reference rebuilds every frame; here it is a no-op unless a transform changed. */
SOLB_API int solb_tlas_regenerate(solb_scene *scene);

/* SURVEY 8f-3 (the reference's "TODO: support multiple instances per BLAS", src/ray/mod.rs:122): one more instance of
 * the BLAS that `source_instance` uses, with its own transform and material.  The new instance id (= gl_InstanceID) is the
 * running count.  Takes effect at the next solb_accel_build / solb_tlas_regenerate. */
SOLB_API int solb_scene_add_instance(solb_scene *scene, uint32_t source_instance, const float transform[16],
                                     uint32_t material_index, uint32_t *out_index);
/* SolbAccelMode for the next build (default SOLB_ACCEL_FLAT). */
SOLB_API int solb_scene_set_accel_mode(solb_scene *scene, uint32_t mode);

SOLB_API int solb_scene_instance_count(solb_scene *scene, uint32_t *out);
/* SceneDescription::instances as the shader sees them (get_instances_buffer, src/ray/mod.rs:186) */
SOLB_API int solb_scene_get_instances(solb_scene *scene, SolbSceneInstance *out, uint32_t capacity);
/* primitive_count of every instance's BLAS (src/ray/acceleration.rs:183), instance order */
SOLB_API int solb_scene_instance_triangles(solb_scene *scene, uint32_t *out, uint32_t capacity);
SOLB_API int solb_accel_info(solb_scene *scene, SolbAccelInfo *out);
/* test/inspection hooks: copy the built structure back (80 B nodes, 48 B triangle records) */
SOLB_API int solb_accel_read_nodes(solb_scene *scene, void *host, size_t bytes);
SOLB_API int solb_accel_read_triangles(solb_scene *scene, void *host, size_t bytes);

/* ---- targets ------------------------------------------------------------------------------- */

/* Replaces create_image_target (examples/5-pathtrace.rs:57-80): zero-initialised. */
SOLB_API int solb_target_create(solb_ctx *ctx, uint32_t width, uint32_t height, uint32_t format, solb_target **out);
SOLB_API int solb_target_destroy(solb_target *t);
SOLB_API int solb_target_clear(solb_target *t);
/* Replaces cmd_blit_to(present image) (examples/5-pathtrace.rs:360-361): D2H of the whole image. Synchronises. */
SOLB_API int solb_target_readback(solb_target *t, void *host, size_t bytes);
SOLB_API int solb_target_upload(solb_target *t, const void *host, size_t bytes);
/* The same copy without the wait, for a render loop with frames in flight (the reference's blit is a queued command too,
 * examples/5-pathtrace.rs:360-361 inside the frame's command buffer): enqueued after the frames traced so far; `host` must
 * stay valid, and is complete, once a fence signalled after this call has been waited for (or after solb_synchronize).
 * The copy overlaps the next frame's tracing when `host` is page-locked (solb_host_alloc). */
SOLB_API int solb_target_readback_async(solb_target *t, void *host, size_t bytes);
SOLB_API int solb_target_device_ptr(solb_target *t, void **out); /* for torch.distributed / NCCL plumbing */
SOLB_API int solb_target_info(solb_target *t, uint32_t *width, uint32_t *height, uint32_t *format);

/* ---- trace launches: each replaces ShaderBindingTable::cmd_trace_rays (src/ray/sbt.rs:167-180)
 *      for one of the three pipelines built by ray::Pipeline::new (src/ray/pipeline.rs:61-122) ---- */

SOLB_API void solb_trace_params_default(SolbTraceParams *p, int pipeline /* 0 pathtrace, 1 ao */);

/* 5-pathtrace: assets/glsl/pathtrace.{rgen,rchit,rmiss}.  accum RGBA32F in/out, render RGBA8 out (may be NULL).
 * Asynchronous: returns once the frame is enqueued (up to three frames of one ctx are in flight; the targets are written
 * in frame order on the ctx stream).  Read results with solb_target_readback (blocks) or readback_async + a fence. */
SOLB_API int solb_trace_pathtrace(solb_scene *scene, const SolbSceneUniforms *uniforms, const SolbTraceParams *params,
                                  solb_target *accum, solb_target *render);
/* 4-ray-ao: assets/glsl/ao.{rgen,rchit,rmiss}.  image RGBA32F in/out; blue_noise = host rgba8[w*h],
 * rows already flipped like Texture2d::new does (src/texture.rs:490-493); uploaded once per pointer change. */
SOLB_API int solb_set_blue_noise(solb_ctx *ctx, const uint8_t *rgba8, uint32_t width, uint32_t height);
SOLB_API int solb_trace_ao(solb_scene *scene, const SolbSceneUniforms *uniforms, const SolbTraceParams *params,
                           solb_target *image);
/* 3-ray-debug: assets/glsl/debug.{rgen,rchit,rmiss}.  render RGBA8 out; ids RG32UI out (optional): the
 * (gl_InstanceID, gl_PrimitiveID) of the primary hit that north_star's parity gate needs;
 * attribs RGBA32F out (optional): (u, v, t, 0). */
SOLB_API int solb_trace_debug(solb_scene *scene, const SolbSceneUniforms *uniforms, solb_target *render,
                              solb_target *ids, solb_target *attribs);
/* traceRayEXT itself (assets/glsl/pathtrace.rgen:65-76) for arbitrary host rays:
 * rays[i] = {ox,oy,oz,tmin, dx,dy,dz,tmax}; hits[i] = {instance, primitive, bits(u), bits(v)}; t optional. */
SOLB_API int solb_trace_rays(solb_scene *scene, const float *rays, uint32_t n, uint32_t *hits, float *t_out);

/* test hook: the builder's onesweep radix sort on host (u64 key, u32 value) pairs, in place */
SOLB_API int solb_test_sort_pairs(solb_ctx *ctx, uint64_t *keys, uint32_t *values, uint32_t n, int key_bits);

/* ---- multi-GPU resolve: out = sum.xyz / sum.w with the reference's display transform ---- */
/* accum_out RGBA32F (may alias sum), render RGBA8 (may be NULL). */
SOLB_API int solb_resolve_sum(solb_ctx *ctx, solb_target *sum, solb_target *accum_out, solb_target *render);

/* ---- multi-GPU exchange (new work: the reference is single-GPU).  One process per GPU, one communicator per
 * ctx; NCCL over NVLink underneath, bound at run time (the copy already loaded in the process, else SOLB_NCCL_LIB, else
 * libnccl.so.2).  All calls are collective over the ranks of the communicator and asynchronous on the ctx stream.
 * solb_allgather_rows moves the bands with peer stores into CUDA-IPC-mapped staging blocks (one process per GPU on one
 * node) and falls back, on every rank together, to ncclAllGather when a peer's block cannot be mapped; SOLB_P2P=0 forces
 * the latter. ---- */
#define SOLB_COMM_ID_BYTES 128
/* rank 0 creates the id and hands the 128 bytes to the other ranks by any host-side means (ncclGetUniqueId) */
SOLB_API int solb_comm_unique_id(uint8_t *id_out /* [SOLB_COMM_ID_BYTES] */);
SOLB_API int solb_comm_init(solb_ctx *ctx, const uint8_t *id /* [SOLB_COMM_ID_BYTES] */, int rank, int world);
SOLB_API int solb_comm_info(solb_ctx *ctx, int *rank, int *world, int *nccl_version);
SOLB_API int solb_comm_destroy(solb_ctx *ctx);
/* Frames (sample) split: rank r renders frames f = r (mod N) with SOLB_ACCUM_SUM into `sum`; this sums the N targets onto
 * `root` in place (ncclReduce, float32) and, on the root, resolves sum / count with the reference's display transform into
 * accum_out (may alias sum) / render (each may be NULL).  Without a communicator (world 1) it is the resolve alone. */
SOLB_API int solb_reduce_accum(solb_ctx *ctx, solb_target *sum, int root, solb_target *accum_out, solb_target *render);
/* Tile split of ONE frame: rank r has traced the bands r, r + N, ... of band_rows rows each (SolbTraceParams tile_row_begin =
 * r * band_rows, tile_row_count = band_rows, tile_row_stride = N * band_rows) into its full-size target; afterwards every
 * rank holds the whole image (pack, ncclAllGather, scatter). */
SOLB_API int solb_allgather_rows(solb_ctx *ctx, solb_target *target, uint32_t band_rows);

#ifdef __cplusplus
}
#endif
#endif /* SOLB_H */
