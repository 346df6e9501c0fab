// host_capi.cpp — flat C exports of the C++ host mirror (sol.hpp) so the Python layer and the CPU test
// tier can drive sol::scene::load_scene / Camera / SceneUniforms without a GPU.
#include <cstdio>
#include <cstring>

#include "sol.hpp"

using namespace sol;

#define SOLH_API extern "C" __attribute__((visibility("default")))

struct SolhMeshInfo {
    const void *vertices; uint32_t n_vertices;
    const uint32_t *indices; uint32_t n_indices;
    uint32_t n_sections;
    float transform[16];
    char name[64];
};

static void set_err(char *err, size_t n, const char *msg) { if (err && n) { std::snprintf(err, n, "%s", msg); } }

SOLH_API scene::Scene *solh_load_scene(const char *path, char *err, size_t errlen) {
    try {
        return new scene::Scene(scene::load_scene(nullptr, path));
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}
SOLH_API void solh_scene_free(scene::Scene *s) { delete s; }
SOLH_API uint32_t solh_scene_mesh_count(const scene::Scene *s) { return (uint32_t)s->meshes.size(); }
SOLH_API uint32_t solh_scene_material_count(const scene::Scene *s) { return (uint32_t)s->materials.size(); }
SOLH_API const void *solh_scene_materials(const scene::Scene *s) { return s->materials.data(); }
SOLH_API int solh_scene_has_camera(const scene::Scene *s) { return s->camera.has_value(); }

SOLH_API uint32_t solh_scene_texture_count(const scene::Scene *s) { return (uint32_t)s->textures.size(); }
// out4 = { width, height, wrap_s, wrap_t }; returns the rgba8 pixels (rows top first), valid until solh_scene_free
SOLH_API const uint8_t *solh_scene_texture(const scene::Scene *s, uint32_t i, uint32_t *out4) {
    if (i >= s->textures.size()) return nullptr;
    const scene::Texture &t = s->textures[i];
    out4[0] = t.width; out4[1] = t.height; out4[2] = t.wrap_s; out4[3] = t.wrap_t;
    return t.rgba8.data();
}
SOLH_API const uint32_t *solh_scene_material_textures(const scene::Scene *s) { return s->material_textures.data(); }
// decode_png for tests: returns 0 and fills w / h; pixels into out when capacity suffices
SOLH_API int solh_decode_png(const uint8_t *data, size_t n, uint32_t *w, uint32_t *h, uint8_t *out, size_t capacity, char *err, size_t errlen) {
    try {
        const image::Rgba8 im = image::decode_png(data, n);
        *w = im.width; *h = im.height;
        if (out && capacity >= im.pixels.size()) std::memcpy(out, im.pixels.data(), im.pixels.size());
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}

SOLH_API int solh_mesh_info(const scene::Scene *s, uint32_t i, SolhMeshInfo *out) {
    if (i >= s->meshes.size()) return -1;
    const scene::Mesh &m = s->meshes[i];
    out->vertices = m.vertices.data(); out->n_vertices = (uint32_t)m.vertices.size();
    out->indices = m.indices.data(); out->n_indices = (uint32_t)m.indices.size();
    out->n_sections = (uint32_t)m.primitive_sections.size();
    std::memcpy(out->transform, m.transform.data(), 64);
    std::snprintf(out->name, sizeof(out->name), "%s", m.name.c_str());
    return 0;
}
SOLH_API uint32_t solh_mesh_extra_instance_count(const scene::Scene *s, uint32_t i) {
    return i < s->meshes.size() ? (uint32_t)s->meshes[i].extra_instance_transforms.size() : 0u;
}
SOLH_API int solh_mesh_extra_instance_transform(const scene::Scene *s, uint32_t i, uint32_t k, float *out) {
    if (i >= s->meshes.size() || k >= s->meshes[i].extra_instance_transforms.size()) return -1;
    std::memcpy(out, s->meshes[i].extra_instance_transforms[k].data(), 64);
    return 0;
}
// material_index = 0xffffffff when the primitive has none; n_indices = 0 when it is not indexed
SOLH_API int solh_mesh_sections(const scene::Scene *s, uint32_t i, SolbSection *out) {
    if (i >= s->meshes.size()) return -1;
    const scene::Mesh &m = s->meshes[i];
    for (size_t k = 0; k < m.primitive_sections.size(); k++) {
        const auto &ps = m.primitive_sections[k];
        out[k].first_vertex = (uint32_t)ps.vertices.offset;
        out[k].n_vertices = (uint32_t)ps.vertices.element_count;
        out[k].first_index = ps.indices ? (uint32_t)ps.indices->offset : 0u;
        out[k].n_indices = ps.indices ? (uint32_t)ps.indices->element_count : 0u;
        out[k].material_index = ps.material_index ? (uint32_t)*ps.material_index : 0xffffffffu;
    }
    return 0;
}

SOLH_API scene::Camera *solh_camera_new(float w, float h) { return new scene::Camera(Vec2{ w, h }); }
SOLH_API scene::Camera *solh_camera_from_scene(const scene::Scene *s) { return s->camera ? new scene::Camera(*s->camera) : nullptr; }
SOLH_API scene::Camera *solh_camera_from_view(const float *view, float yfov, float zn, float zf) {
    Mat4 v;
    std::memcpy(v.data(), view, 64);
    return new scene::Camera(scene::Camera::from_view(v, yfov, zn, zf));
}
SOLH_API void solh_camera_free(scene::Camera *c) { delete c; }
SOLH_API void solh_camera_look_at(scene::Camera *c, const float *eye, const float *center, const float *up) {
    c->look_at({ eye[0], eye[1], eye[2] }, { center[0], center[1], center[2] }, { up[0], up[1], up[2] });
}
SOLH_API void solh_camera_set_window_size(scene::Camera *c, float w, float h) { c->set_window_size({ w, h }); }
SOLH_API void solh_camera_set_vfov(scene::Camera *c, float vfov) { c->set_vfov(vfov); }
SOLH_API void solh_camera_matrices(const scene::Camera *c, float *view, float *persp) {
    std::memcpy(view, c->view_matrix().data(), 64);
    std::memcpy(persp, c->perspective_matrix().data(), 64);
}
SOLH_API void solh_scene_uniforms(const scene::Camera *c, uint32_t w, uint32_t h, uint32_t frame, SolbSceneUniforms *out) {
    const SceneUniforms u = SceneUniforms::from(*c, { w, h, frame });
    std::memcpy(out, &u, sizeof(SolbSceneUniforms));
}
SOLH_API void solh_mat4_inverse(const float *m, float *out) {
    Mat4 a;
    std::memcpy(a.data(), m, 64);
    const Mat4 r = math::inverse(a);
    std::memcpy(out, r.data(), 64);
}
SOLH_API void solh_mat4_mul(const float *a, const float *b, float *out) {
    Mat4 x, y;
    std::memcpy(x.data(), a, 64);
    std::memcpy(y.data(), b, 64);
    const Mat4 r = math::mul(x, y);
    std::memcpy(out, r.data(), 64);
}
SOLH_API void solh_mat4_from_scale_rotation_x(float scale, float angle, float *out) {
    // examples/4-ray-ao.rs:84-85: Mat4::from_scale(splat(s)) * Mat4::from_rotation_x(a)
    const Mat4 r = math::mul(math::from_scale({ scale, scale, scale }), math::from_rotation_x(angle));
    std::memcpy(out, r.data(), 64);
}
