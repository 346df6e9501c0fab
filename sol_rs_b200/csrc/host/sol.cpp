// sol.cpp — host-side mirror implementation: math (glam restatement), Context/Image2d wrappers,
// scene::Camera, SceneUniforms, ray::{SceneDescription, Pipeline, ShaderBindingTable}.
#include "sol.hpp"

#include <sys/stat.h>

#include <algorithm>

namespace sol {

// ---- math ------------------------------------------------------------------------------------------------
namespace math {

Mat4 identity() {
    Mat4 m{};
    m[0] = m[5] = m[10] = m[15] = 1.0f;
    return m;
}

Mat4 mul(const Mat4 &a, const Mat4 &b) {
    Mat4 r{};
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            float acc = 0.0f;
            for (int k = 0; k < 4; k++) acc = acc + a[4 * k + row] * b[4 * c + k];
            r[4 * c + row] = acc;
        }
    return r;
}

Mat4 transpose(const Mat4 &m) {
    Mat4 r{};
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) r[4 * c + row] = m[4 * row + c];
    return r;
}

// glam Mat4::inverse, scalar-math path (GLM cofactor scheme)
Mat4 inverse(const Mat4 &m) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3], m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    const float m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11], m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];
    const float c00 = m22 * m33 - m32 * m23, c02 = m12 * m33 - m32 * m13, c03 = m12 * m23 - m22 * m13;
    const float c04 = m21 * m33 - m31 * m23, c06 = m11 * m33 - m31 * m13, c07 = m11 * m23 - m21 * m13;
    const float c08 = m21 * m32 - m31 * m22, c10 = m11 * m32 - m31 * m12, c11 = m11 * m22 - m21 * m12;
    const float c12 = m20 * m33 - m30 * m23, c14 = m10 * m33 - m30 * m13, c15 = m10 * m23 - m20 * m13;
    const float c16 = m20 * m32 - m30 * m22, c18 = m10 * m32 - m30 * m12, c19 = m10 * m22 - m20 * m12;
    const float c20 = m20 * m31 - m30 * m21, c22 = m10 * m31 - m30 * m11, c23 = m10 * m21 - m20 * m11;
    const float f0[4] = { c00, c00, c02, c03 }, f1[4] = { c04, c04, c06, c07 }, f2[4] = { c08, c08, c10, c11 };
    const float f3[4] = { c12, c12, c14, c15 }, f4[4] = { c16, c16, c18, c19 }, f5[4] = { c20, c20, c22, c23 };
    const float v0[4] = { m10, m00, m00, m00 }, v1[4] = { m11, m01, m01, m01 }, v2[4] = { m12, m02, m02, m02 }, v3[4] = { m13, m03, m03, m03 };
    const float sa[4] = { 1, -1, 1, -1 }, sb[4] = { -1, 1, -1, 1 };
    Mat4 inv{};
    for (int i = 0; i < 4; i++) {
        inv[0 + i] = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa[i];
        inv[4 + i] = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb[i];
        inv[8 + i] = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa[i];
        inv[12 + i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb[i];
    }
    const float dot1 = ((m00 * inv[0] + m01 * inv[4]) + m02 * inv[8]) + m03 * inv[12];
    const float rcp = 1.0f / dot1;
    for (auto &x : inv) x *= rcp;
    return inv;
}

// glam Mat4::perspective_rh: right-handed, depth 0..1, no Y flip (src/scene/camera.rs:104-105)
Mat4 perspective_rh(float fov_y_radians, float aspect, float z_near, float z_far) {
    const float s = sinf(0.5f * fov_y_radians), c = cosf(0.5f * fov_y_radians);
    const float h = c / s, w = h / aspect, r = z_far / (z_near - z_far);
    Mat4 m{};
    m[0] = w;
    m[5] = h;
    m[10] = r; m[11] = -1.0f;
    m[14] = r * z_near;
    return m;
}

static Vec3 normalize(Vec3 v) {
    const float rl = 1.0f / sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
    return { v.x * rl, v.y * rl, v.z * rl };
}
static Vec3 cross(Vec3 a, Vec3 b) { return { a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y }; }
static float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// glam Mat4::look_at_rh(eye, center, up) == look_to_lh(eye, eye - center, up) (src/scene/camera.rs:99)
Mat4 look_at_rh(Vec3 eye, Vec3 center, Vec3 up) {
    const Vec3 f = normalize({ eye.x - center.x, eye.y - center.y, eye.z - center.z });
    const Vec3 s = normalize(cross(up, f));
    const Vec3 u = cross(f, s);
    Mat4 m{};
    m[0] = s.x; m[1] = u.x; m[2] = f.x;
    m[4] = s.y; m[5] = u.y; m[6] = f.y;
    m[8] = s.z; m[9] = u.z; m[10] = f.z;
    m[12] = -dot(s, eye); m[13] = -dot(u, eye); m[14] = -dot(f, eye); m[15] = 1.0f;
    return m;
}

Mat4 from_scale(Vec3 s) {
    Mat4 m{};
    m[0] = s.x; m[5] = s.y; m[10] = s.z; m[15] = 1.0f;
    return m;
}

Mat4 from_rotation_x(float angle) {
    const float s = sinf(angle), c = cosf(angle);
    Mat4 m = identity();
    m[5] = c; m[6] = s;
    m[9] = -s; m[10] = c;
    return m;
}

}  // namespace math

// ---- Context / Image2d ---------------------------------------------------------------------------------
std::shared_ptr<Context> Context::create(int device, void *cuda_stream) {
    std::shared_ptr<Context> c(new Context());
    const int rc = solb_ctx_create(device, cuda_stream, &c->h_);
    if (rc != SOLB_OK) throw Error(rc, solb_last_error(nullptr));
    return c;
}
Context::~Context() { solb_ctx_destroy(h_); }
void Context::check(int rc) const {
    if (rc != SOLB_OK) throw Error(rc, solb_last_error(h_));
}

Image2d::Image2d(std::shared_ptr<Context> ctx, uint32_t width, uint32_t height, SolbTargetFormat format)
    : ctx_(std::move(ctx)), w_(width), hgt_(height), fmt_(format) {
    ctx_->check(solb_target_create(ctx_->handle(), width, height, (uint32_t)format, &h_));
}
Image2d::~Image2d() { solb_target_destroy(h_); }
size_t Image2d::size_bytes() const {
    return (size_t)w_ * hgt_ * (fmt_ == SOLB_FORMAT_RGBA32F ? 16 : (fmt_ == SOLB_FORMAT_RGBA8 ? 4 : 8));
}

// ---- Camera ----------------------------------------------------------------------------------------------
namespace scene {

Camera::Camera(Vec2 window_size) : window_size_(window_size) { update_persp(); }  // camera.rs:54-71

Camera Camera::from_view(const Mat4 &view, float yfov, float z_near, float z_far) {  // camera.rs:73-94
    Camera c;
    const Mat4 vi = math::inverse(view);
    c.position_ = { vi[12], vi[13], vi[14] };
    c.up_ = { vi[4], vi[5], vi[6] };
    c.center_ = { c.position_.x + vi[8] * -4.0f, c.position_.y + vi[9] * -4.0f, c.position_.z + vi[10] * -4.0f };
    c.vfov_ = yfov;
    c.z_near_ = z_near;
    c.z_far_ = z_far;
    c.view_matrix_ = view;
    c.persp_matrix_ = math::identity();
    c.window_size_ = { 1920.0f, 1080.0f };
    return c;
}
void Camera::update_view() { view_matrix_ = math::look_at_rh(position_, center_, up_); }
void Camera::update_persp() {
    const float aspect = window_size_.x / window_size_.y;
    persp_matrix_ = math::perspective_rh(math::to_radians(vfov_), aspect, z_near_, z_far_);  // camera.rs:102-106
}
void Camera::look_at(Vec3 eye, Vec3 center, Vec3 up) { position_ = eye; center_ = center; up_ = up; update_view(); }
void Camera::set_window_size(Vec2 ws) { window_size_ = ws; update_persp(); }
void Camera::set_vfov(float vfov) { vfov_ = vfov; update_persp(); }

}  // namespace scene

SceneUniforms SceneUniforms::from(const scene::Camera &camera, UVec3 frame) {  // examples/5-pathtrace.rs:19-31
    SceneUniforms u;
    std::memset(&u, 0, sizeof(u));
    const Mat4 vp = math::mul(camera.perspective_matrix(), camera.view_matrix());
    const Mat4 id = math::identity(), vi = math::inverse(camera.view_matrix()), pi = math::inverse(camera.perspective_matrix());
    std::memcpy(u.model, id.data(), 64);
    std::memcpy(u.view, camera.view_matrix().data(), 64);
    std::memcpy(u.view_inverse, vi.data(), 64);
    std::memcpy(u.projection, camera.perspective_matrix().data(), 64);
    std::memcpy(u.projection_inverse, pi.data(), 64);
    std::memcpy(u.model_view_projection, vp.data(), 64);
    u.frame[0] = frame.x; u.frame[1] = frame.y; u.frame[2] = frame.z;
    return u;
}

// ---- ray -------------------------------------------------------------------------------------------------
namespace ray {

SceneDescription SceneDescription::from_scene(std::shared_ptr<Context> context, const scene::Scene &scene, SolbAccelMode accel_mode,
                                              bool instancing) {  // src/ray/mod.rs:50-57
    std::vector<const scene::Mesh *> meshes;
    std::vector<Mat4> transforms;
    for (const auto &m : scene.meshes) { meshes.push_back(&m); transforms.push_back(m.transform); }
    SceneDescription sd = from_meshes(std::move(context), meshes, transforms, &scene.materials);
    bool rebuild = accel_mode != SOLB_ACCEL_FLAT;
    if (rebuild) sd.set_accel_mode(accel_mode);
    if (instancing) {
        // instance ids run over meshes x sections (src/ray/mod.rs:113); further nodes of a mesh append instances of the same BLASes
        size_t first_instance = 0;
        for (const auto &m : scene.meshes) {
            for (const Mat4 &t : m.extra_instance_transforms)
                for (size_t k = 0; k < m.primitive_sections.size(); k++) {
                    sd.add_instance(first_instance + k, t, (uint32_t)*m.primitive_sections[k].material_index);
                    rebuild = true;
                }
            first_instance += m.primitive_sections.size();
        }
    }
    if (rebuild) sd.accel_build();
    return sd;
}

SceneDescription SceneDescription::from_meshes(std::shared_ptr<Context> context, const std::vector<const scene::Mesh *> &meshes,
                                               const std::vector<Mat4> &mesh_transforms,
                                               const std::vector<scene::MaterialInfo> *materials) {  // src/ray/mod.rs:59-156
    if (meshes.size() != mesh_transforms.size()) throw Error(SOLB_ERR_INVALID, "from_meshes: meshes/transforms length mismatch");
    std::vector<SolbMeshDesc> descs(meshes.size());
    std::vector<std::vector<SolbSection>> sections(meshes.size());
    for (size_t i = 0; i < meshes.size(); i++) {
        const scene::Mesh &m = *meshes[i];
        for (const auto &ps : m.primitive_sections) {
            SolbSection s;
            s.first_vertex = (uint32_t)ps.vertices.offset;
            s.n_vertices = (uint32_t)ps.vertices.element_count;
            s.first_index = ps.indices ? (uint32_t)ps.indices->offset : 0u;
            s.n_indices = ps.indices ? (uint32_t)ps.indices->element_count : 0u;
            // the reference unwrap()s material_index when a material buffer is given (src/scene/mod.rs:65)
            if (!ps.material_index) throw Error(SOLB_ERR_INVALID, "primitive without material (Option::unwrap on None)");
            s.material_index = (uint32_t)*ps.material_index;
            sections[i].push_back(s);
        }
        SolbMeshDesc &d = descs[i];
        d.vertices = m.vertices.data();
        d.n_vertices = (uint32_t)m.vertices.size();
        d.indices = m.indices.data();
        d.n_indices = (uint32_t)m.indices.size();
        d.sections = sections[i].data();
        d.n_sections = (uint32_t)sections[i].size();
        std::memcpy(d.transform, mesh_transforms[i].data(), 64);
    }
    SceneDescription sd;
    sd.ctx_ = std::move(context);
    sd.ctx_->check(solb_scene_create(sd.ctx_->handle(), descs.data(), (uint32_t)descs.size(), materials ? materials->data() : nullptr,
                                     materials ? (uint32_t)materials->size() : 0u, &sd.h_));
    const int rc = solb_accel_build(sd.h_);  // BLAS::new per section + TLAS::new + end_single_time_cmd
    if (rc != SOLB_OK) { const std::string msg = solb_last_error(sd.ctx_->handle()); throw Error(rc, msg); }
    return sd;
}

SceneDescription::SceneDescription(SceneDescription &&o) noexcept : ctx_(std::move(o.ctx_)), h_(o.h_) { o.h_ = nullptr; }
SceneDescription::~SceneDescription() { if (h_) solb_scene_destroy(h_); }

void SceneDescription::blas_transform(const Mat4 &transform, size_t index) {
    ctx_->check(solb_instance_set_transform(h_, (uint32_t)index, transform.data()));
}
void SceneDescription::blas_transforms(const std::vector<Mat4> &transforms) {
    for (size_t i = 0; i < transforms.size(); i++) blas_transform(transforms[i], i);
}
uint32_t SceneDescription::add_instance(size_t source_instance, const Mat4 &transform, uint32_t material_index) {
    uint32_t id = 0;
    ctx_->check(solb_scene_add_instance(h_, (uint32_t)source_instance, transform.data(), material_index, &id));
    return id;
}
void SceneDescription::set_textures(const std::vector<scene::Texture> &textures, const std::vector<uint32_t> &material_textures) {
    std::vector<SolbTextureDesc> descs(textures.size());
    for (size_t i = 0; i < textures.size(); i++) {
        if (textures[i].rgba8.size() != (size_t)textures[i].width * textures[i].height * 4)
            throw Error(SOLB_ERR_INVALID, "set_textures: texture size does not match its pixels");
        descs[i] = SolbTextureDesc{ textures[i].rgba8.data(), textures[i].width, textures[i].height, textures[i].wrap_s, textures[i].wrap_t,
                                    1u /* glTF base colour is sRGB-encoded */, 0u };
    }
    ctx_->check(solb_scene_set_textures(h_, descs.data(), (uint32_t)descs.size(), material_textures.data(), (uint32_t)material_textures.size()));
}
void SceneDescription::set_accel_mode(SolbAccelMode mode) { ctx_->check(solb_scene_set_accel_mode(h_, (uint32_t)mode)); }
void SceneDescription::accel_build() { ctx_->check(solb_accel_build(h_)); }
void SceneDescription::tlas_regenerate() { ctx_->check(solb_tlas_regenerate(h_)); }
void SceneDescription::update() { ctx_->check(solb_scene_update(h_)); }
size_t SceneDescription::blas_count() const {
    uint32_t n = 0;
    ctx_->check(solb_scene_instance_count(h_, &n));
    return n;
}
std::vector<SceneInstance> SceneDescription::instances() const {
    std::vector<SceneInstance> v(blas_count());
    ctx_->check(solb_scene_get_instances(h_, v.data(), (uint32_t)v.size()));
    return v;
}
SolbAccelInfo SceneDescription::accel_info() const {
    SolbAccelInfo i;
    ctx_->check(solb_accel_info(h_, &i));
    return i;
}

static bool ends_with(const std::string &s, const std::string &suffix) {
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

Pipeline::Pipeline(std::shared_ptr<Context> context, const PipelineInfo &info) {
    bool have_rgen = false;
    for (const auto &sh : info.shaders_) {
        if (sh.second != ShaderStage::RAYGEN_KHR) continue;
        have_rgen = true;
        if (ends_with(sh.first, "pathtrace.rgen")) kind_ = PipelineKind::PATHTRACE;
        else if (ends_with(sh.first, "ao.rgen")) kind_ = PipelineKind::AO;
        else if (ends_with(sh.first, "debug.rgen")) kind_ = PipelineKind::DEBUG;
        else throw Error(SOLB_ERR_UNSUPPORTED, "ray::Pipeline: no CUDA kernel family for raygen shader '" + sh.first + "'");
    }
    if (!have_rgen) throw Error(SOLB_ERR_INVALID, "ray::Pipeline: no raygen stage");
    // specialization constant 0 = ENABLE_SKYLIGHT (examples/5-pathtrace.rs:103, pathtrace.rmiss:5)
    enable_sky_ = info.spec_id_ == 0 && !info.spec_.empty() && info.spec_[0] != 0;
    // the reference compiles its GLSL stages here (src/ray/pipeline.rs:105-115); the counterpart is loading the kernels
    if (context) {
        const int rc = solb_ctx_preload(context->handle());
        if (rc != SOLB_OK) throw Error(rc, std::string("ray::Pipeline: ") + solb_last_error(context->handle()));
    }
}

ShaderBindingTable::ShaderBindingTable(std::shared_ptr<Context> context, const Pipeline &pipeline, const ShaderBindingTableInfo &info)
    : ctx_(std::move(context)), kind_(pipeline.kind()), enable_sky_(pipeline.enable_sky()) {
    if (info.raygen_.size() != 1 || info.miss_.size() != 1 || info.hit_.size() != 1)
        throw Error(SOLB_ERR_UNSUPPORTED, "ShaderBindingTable: exactly one raygen / miss / hit group is supported");
}

void ShaderBindingTable::cmd_trace_rays(const TraceBindings &b, Extent3D extent) const {
    if (!b.scene_description || !b.uniforms) throw Error(SOLB_ERR_INVALID, "cmd_trace_rays: scene description / uniforms not bound");
    Image2d *first = b.accum_target ? b.accum_target : b.render_target;
    if (!first) throw Error(SOLB_ERR_INVALID, "cmd_trace_rays: no storage image bound");
    if (extent.width != first->width() || extent.height != first->height() || extent.depth != 1)
        throw Error(SOLB_ERR_INVALID, "cmd_trace_rays: extent differs from the bound targets");
    SolbTraceParams p;
    solb_trace_params_default(&p, kind_ == PipelineKind::AO ? 1 : 0);
    p.accum_start_frame = (int32_t)b.accumulation_start_frame;
    p.enable_sky = enable_sky_ ? 1u : 0u;
    if (b.overrides.samples_per_frame) p.samples_per_frame = b.overrides.samples_per_frame;
    if (b.overrides.max_bounces) p.max_bounces = b.overrides.max_bounces;
    if (b.overrides.schedule) p.schedule = b.overrides.schedule;  // 0 keeps the default (AUTO)
    p.accum_mode = b.overrides.accum_mode;
    p.collect_stats = b.overrides.collect_stats;
    p.tile_row_begin = b.overrides.tile_row_begin;
    p.tile_row_count = b.overrides.tile_row_count;
    p.tile_row_stride = b.overrides.tile_row_stride;
    solb_scene *s = b.scene_description->handle();
    switch (kind_) {
        case PipelineKind::PATHTRACE:
            if (!b.accum_target) throw Error(SOLB_ERR_INVALID, "pathtrace: accumulation image not bound");
            ctx_->check(solb_trace_pathtrace(s, b.uniforms, &p, b.accum_target->handle(), b.render_target ? b.render_target->handle() : nullptr));
            break;
        case PipelineKind::AO:
            if (!b.accum_target) throw Error(SOLB_ERR_INVALID, "ao: image not bound");
            ctx_->check(solb_trace_ao(s, b.uniforms, &p, b.accum_target->handle()));
            break;
        case PipelineKind::DEBUG:
            ctx_->check(solb_trace_debug(s, b.uniforms, b.render_target ? b.render_target->handle() : nullptr,
                                         b.ids_target ? b.ids_target->handle() : nullptr, nullptr));
            break;
    }
}

}  // namespace ray

namespace util {
std::optional<std::string> find_asset(const std::string &relative, const std::string &start) {
    std::string dir = start.empty() ? std::string(".") : start;
    for (int i = 0; i < 6; i++) {
        const std::string cand = dir + "/assets/" + relative;
        struct stat st;
        if (stat(cand.c_str(), &st) == 0) return cand;
        dir += "/..";
    }
    return std::nullopt;
}
}  // namespace util

}  // namespace sol
