// sol.hpp — C++ host-side mirror of the reference's Rust API for the ray-tracing path.
//
// The reference toolchain (rustc/cargo) is absent from this image, so the host layer that sits above
// the C ABI (include/solb.h) is written in C++ with the same names, argument meaning and error behaviour
// as the Rust items it mirrors (paths relative to the reference root):
//   sol::scene::{load_scene, Scene, Mesh, ModelVertex, MaterialInfo, PrimitiveSection, Camera}
//        <- src/scene/mod.rs, src/scene/mesh.rs, src/scene/camera.rs
//   sol::ray::{SceneDescription, SceneInstance, PipelineInfo, Pipeline, ShaderBindingTableInfo, ShaderBindingTable}
//        <- src/ray/mod.rs, src/ray/pipeline.rs, src/ray/sbt.rs
//   sol::SceneUniforms  <- examples/5-pathtrace.rs:7-32
//   sol::Context / sol::Image2d <- src/context.rs, src/texture.rs (only what this path needs)
// Errors: the reference unwrap()s / panics on this path; here a failed C-ABI call throws sol::Error.
#pragma once
#include <array>
#include <cmath>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/solb.h"

#pragma GCC visibility push(default)  // the C++ mirror is libsol_host.so's public API
namespace sol {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

namespace image {
struct Rgba8 {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> pixels;  // rows top first
};
bool is_png(const uint8_t *data, size_t n);
Rgba8 decode_png(const uint8_t *data, size_t n);  // host/png.cpp; throws sol::Error
}  // namespace image

using Mat4 = std::array<float, 16>;  // column-major like glam::Mat4
struct Vec3 { float x, y, z; };
struct Vec2 { float x, y; };
struct UVec3 { uint32_t x, y, z; };
struct Extent3D { uint32_t width, height, depth; };

// ---- glam 0.20.2 restatements (Cargo.lock:509-510), f32 ---------------------------------------------
namespace math {
Mat4 identity();
Mat4 mul(const Mat4 &a, const Mat4 &b);
Mat4 inverse(const Mat4 &m);
Mat4 transpose(const Mat4 &m);
Mat4 perspective_rh(float fov_y_radians, float aspect, float z_near, float z_far);
Mat4 look_at_rh(Vec3 eye, Vec3 center, Vec3 up);
Mat4 from_scale(Vec3 s);
Mat4 from_rotation_x(float angle);
inline float to_radians(float deg) { return deg * (3.14159265358979323846f / 180.0f); }
}  // namespace math

// ---- Context / Image2d -------------------------------------------------------------------------------
class Context {
public:
    static std::shared_ptr<Context> create(int device = 0, void *cuda_stream = nullptr);
    ~Context();
    solb_ctx *handle() const { return h_; }
    void check(int rc) const;
    void synchronize() const { check(solb_synchronize(h_)); }
    SolbStats stats() const { SolbStats s; check(solb_stats_get(h_, &s)); return s; }

private:
    Context() = default;
    solb_ctx *h_ = nullptr;
};

class Image2d {
public:
    Image2d(std::shared_ptr<Context> ctx, uint32_t width, uint32_t height, SolbTargetFormat format);
    ~Image2d();
    Image2d(const Image2d &) = delete;
    Image2d &operator=(const Image2d &) = delete;
    solb_target *handle() const { return h_; }
    uint32_t width() const { return w_; }
    uint32_t height() const { return hgt_; }
    size_t size_bytes() const;
    void clear() { ctx_->check(solb_target_clear(h_)); }
    void readback(void *host, size_t bytes) { ctx_->check(solb_target_readback(h_, host, bytes)); }

private:
    std::shared_ptr<Context> ctx_;
    solb_target *h_ = nullptr;
    uint32_t w_ = 0, hgt_ = 0;
    SolbTargetFormat fmt_;
};

// ---- scene ---------------------------------------------------------------------------------------------
namespace scene {

using ModelVertex = SolbModelVertex;    // src/scene/mesh.rs:9-14
using MaterialInfo = SolbMaterialInfo;  // src/scene/mod.rs:19-29

struct BufferPart { size_t offset = 0, element_count = 0; };  // src/scene/mod.rs:31-35

struct PrimitiveSection {  // src/scene/mod.rs:37-44
    size_t index = 0;
    BufferPart vertices;
    std::optional<BufferPart> indices;
    std::optional<size_t> material_index;
    uint32_t get_vertex_count() const { return (uint32_t)vertices.element_count; }
    uint32_t get_vertex_offset() const { return (uint32_t)vertices.offset; }
    uint32_t get_index_count() const { return (uint32_t)indices.value().element_count; }  // unwrap()
};

struct Mesh {  // src/scene/mesh.rs:53-61 (buffers kept as host arrays; SceneDescription uploads them)
    std::string name;
    std::vector<ModelVertex> vertices;
    std::vector<uint32_t> indices;
    Mat4 transform;
    std::vector<PrimitiveSection> primitive_sections;
    // Beyond the reference (glTF node-graph instancing): global transforms of the OTHER nodes that reference
    // this mesh, in node-index order.  The reference gives a mesh only the first node's transform (`transform`, found by
    // its own search, src/scene/mod.rs:106-136); SceneDescription::from_scene(..., instancing = true) turns these into
    // additional instances of the same BLASes.
    std::vector<Mat4> extra_instance_transforms;
};

class Camera {  // src/scene/camera.rs:34-127,264-270 (mouse manipulators are out of scope)
public:
    Camera() = default;
    explicit Camera(Vec2 window_size);
    static Camera from_view(const Mat4 &view, float yfov, float z_near, float z_far);
    void look_at(Vec3 eye, Vec3 center, Vec3 up);
    void set_window_size(Vec2 window_size);
    void set_vfov(float vfov);
    const Mat4 &view_matrix() const { return view_matrix_; }
    const Mat4 &perspective_matrix() const { return persp_matrix_; }
    Vec3 position() const { return position_; }

private:
    void update_view();
    void update_persp();
    Vec3 position_{ 10, 10, 10 }, center_{ 0, 0, 0 }, up_{ 0, -1, 0 };
    float vfov_ = 35.0f, z_near_ = 0.1f, z_far_ = 1000.0f;
    Mat4 view_matrix_ = math::identity(), persp_matrix_ = math::identity();
    Vec2 window_size_{ 1920.0f, 1080.0f };
};

// Base-colour texture of a glTF material (beyond the reference, whose load_scene reads no images): decoded
// rgba8, rows top first, with the sampler's wrap modes (glTF codes; 10497 = REPEAT is the default).
struct Texture {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba8;
    uint32_t wrap_s = 10497, wrap_t = 10497;
};

struct Scene {  // src/scene/mod.rs:99-104
    std::vector<Mesh> meshes;
    std::vector<MaterialInfo> materials;
    std::optional<Camera> camera;
    // beyond the reference: the base-colour textures the materials name (PNG images only; a material whose image cannot be
    // decoded, or that uses a texture-coordinate set other than 0, stays untextured) and, per material, its texture or
    // SOLB_NO_TEXTURE.  SceneDescription::set_textures(scene) binds them; nothing samples them otherwise.
    std::vector<Texture> textures;
    std::vector<uint32_t> material_textures;
};

// src/scene/mod.rs:138-295.  Throws sol::Error where the reference unwrap()s (missing file, bad glTF).
Scene load_scene(std::shared_ptr<Context> context, const std::string &filepath);

}  // namespace scene

// examples/5-pathtrace.rs:7-32
struct SceneUniforms : SolbSceneUniforms {
    static SceneUniforms from(const scene::Camera &camera, UVec3 frame);
};

// ---- ray -----------------------------------------------------------------------------------------------
namespace ray {

using SceneInstance = SolbSceneInstance;  // src/ray/mod.rs:16-24

class SceneDescription {  // src/ray/mod.rs:38-206
public:
    // instancing: every further node referencing a mesh becomes one more instance of its BLASes (accel_mode is then best
    // SOLB_ACCEL_TWO_LEVEL so they share geometry)
    static SceneDescription from_scene(std::shared_ptr<Context> context, const scene::Scene &scene,
                                       SolbAccelMode accel_mode = SOLB_ACCEL_FLAT, bool instancing = false);
    static SceneDescription from_meshes(std::shared_ptr<Context> context, const std::vector<const scene::Mesh *> &meshes,
                                        const std::vector<Mat4> &mesh_transforms, const std::vector<scene::MaterialInfo> *materials);
    SceneDescription(SceneDescription &&o) noexcept;
    SceneDescription &operator=(SceneDescription &&) = delete;
    ~SceneDescription();
    // NOTE (src/ray/mod.rs:133): the reference keys blas_to_instances by MESH index; here `index` is the
    // BLAS == instance index, which is what the name says and what single-section meshes make identical.
    void blas_transform(const Mat4 &transform, size_t index);
    void blas_transforms(const std::vector<Mat4> &transforms);
    // Beyond the reference (its TODO at src/ray/mod.rs:122): one more instance of the BLAS `source_instance` uses; returns the
    // new gl_InstanceID.  accel_mode SOLB_ACCEL_TWO_LEVEL keeps BLASes shared and lets tlas_regenerate rebuild the TLAS only.
    uint32_t add_instance(size_t source_instance, const Mat4 &transform, uint32_t material_index);
    // Beyond the reference: bind the scene's base-colour textures (solb_scene_set_textures); every instance's
    // texture_offset (src/ray/mod.rs:20) becomes the texture of its material.  An empty list unbinds.
    void set_textures(const std::vector<scene::Texture> &textures, const std::vector<uint32_t> &material_textures);
    void set_textures(const scene::Scene &scene) { set_textures(scene.textures, scene.material_textures); }
    void set_accel_mode(SolbAccelMode mode);
    void accel_build();
    void tlas_regenerate();  // the reference takes the command buffer; launches are stream-ordered here
    void update();
    std::vector<SceneInstance> instances() const;
    size_t blas_count() const;
    SolbAccelInfo accel_info() const;
    solb_scene *handle() const { return h_; }
    const std::shared_ptr<Context> &context() const { return ctx_; }

private:
    SceneDescription() = default;
    std::shared_ptr<Context> ctx_;
    solb_scene *h_ = nullptr;
};

enum class ShaderStage { RAYGEN_KHR, MISS_KHR, CLOSEST_HIT_KHR };

class PipelineInfo {  // src/ray/pipeline.rs:5-52
public:
    PipelineInfo &shader(const std::string &path, ShaderStage stage) { shaders_.push_back({ path, stage }); return *this; }
    PipelineInfo &specialization(const uint32_t *data, size_t count, uint32_t constant_id) {
        spec_.assign(data, data + count); spec_id_ = constant_id; return *this;
    }
    PipelineInfo &name(const std::string &n) { name_ = n; return *this; }
    std::vector<std::pair<std::string, ShaderStage>> shaders_;
    std::vector<uint32_t> spec_;
    uint32_t spec_id_ = 0;
    std::string name_;
};

enum class PipelineKind { PATHTRACE, AO, DEBUG };

// ray::Pipeline::new (src/ray/pipeline.rs:61-122): the reference compiles the three GLSL stages at run
// time; the CUDA equivalents are compiled ahead of time, so "creating" a pipeline selects the kernel
// family from the raygen shader's file name and captures specialization constant 0 (ENABLE_SKYLIGHT).
class Pipeline {
public:
    Pipeline(std::shared_ptr<Context> context, const PipelineInfo &info);
    PipelineKind kind() const { return kind_; }
    bool enable_sky() const { return enable_sky_; }

private:
    PipelineKind kind_ = PipelineKind::PATHTRACE;
    bool enable_sky_ = false;
};

class ShaderBindingTableInfo {  // src/ray/sbt.rs:13-55
public:
    ShaderBindingTableInfo &raygen(uint32_t i) { raygen_.push_back(i); return *this; }
    ShaderBindingTableInfo &miss(uint32_t i) { miss_.push_back(i); return *this; }
    ShaderBindingTableInfo &hitgroup(uint32_t i) { hit_.push_back(i); return *this; }
    std::vector<uint32_t> raygen_, miss_, hit_;
};

// What the reference binds through descriptor sets + push constants before cmd_trace_rays
// (examples/5-pathtrace.rs:297-356): set 0 binding 0 uniforms; set 1: TLAS, accum image, render image.
struct TraceBindings {
    const SceneDescription *scene_description = nullptr;
    const SceneUniforms *uniforms = nullptr;
    uint32_t accumulation_start_frame = 0;  // push constant
    Image2d *accum_target = nullptr;        // set 1 binding 1 (pathtrace: accum, ao: image)
    Image2d *render_target = nullptr;       // set 1 binding 2 (pathtrace) / binding 1 (debug)
    Image2d *ids_target = nullptr;          // debug only: (instance, primitive) ids
    SolbTraceParams overrides;              // samples_per_frame / max_bounces / schedule; 0 = reference literal
    TraceBindings() { std::memset(&overrides, 0, sizeof(overrides)); }
};

class ShaderBindingTable {  // src/ray/sbt.rs:58-181
public:
    ShaderBindingTable(std::shared_ptr<Context> context, const Pipeline &pipeline, const ShaderBindingTableInfo &info);
    // src/ray/sbt.rs:167-180.  extent must match the bound targets (the reference passes the window extent).
    void cmd_trace_rays(const TraceBindings &bindings, Extent3D extent) const;

private:
    std::shared_ptr<Context> ctx_;
    PipelineKind kind_;
    bool enable_sky_;
};

}  // namespace ray

namespace util {
std::optional<std::string> find_asset(const std::string &relative, const std::string &start = "");  // src/util.rs:13-31
}

}  // namespace sol
#pragma GCC visibility pop
