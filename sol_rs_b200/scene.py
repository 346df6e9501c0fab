"""sol::scene mirror: load_scene, Scene, Mesh, PrimitiveSection, Camera, SceneUniforms.

The work is done by the C++ host library (csrc/host/{gltf,sol}.cpp -> libsol_host.so), which mirrors
/root/reference src/scene/mod.rs:106-295 and src/scene/camera.rs:54-127; this module only wraps it."""
import ctypes
import os

import numpy as np

from . import _native as N

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libsol_host.so")
_H = None


class _MeshInfo(ctypes.Structure):
    _fields_ = [("vertices", ctypes.c_void_p), ("n_vertices", ctypes.c_uint32), ("indices", ctypes.c_void_p),
                ("n_indices", ctypes.c_uint32), ("n_sections", ctypes.c_uint32), ("transform", ctypes.c_float * 16),
                ("name", ctypes.c_char * 64)]


def hostlib():
    global _H
    if _H is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise N.SolbError(-2, "%s not found: run __graft_entry__.build()" % HOST_LIB_PATH)
        N.lib()  # libsol_host links libsolb
        H = ctypes.CDLL(HOST_LIB_PATH)
        vp, u32, f = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float
        H.solh_load_scene.restype = vp
        H.solh_load_scene.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        H.solh_scene_free.argtypes = [vp]
        H.solh_scene_mesh_count.restype = u32
        H.solh_scene_mesh_count.argtypes = [vp]
        H.solh_scene_material_count.restype = u32
        H.solh_scene_material_count.argtypes = [vp]
        H.solh_scene_materials.restype = vp
        H.solh_scene_materials.argtypes = [vp]
        H.solh_scene_has_camera.argtypes = [vp]
        H.solh_scene_texture_count.restype = u32
        H.solh_scene_texture_count.argtypes = [vp]
        H.solh_scene_texture.restype = vp
        H.solh_scene_texture.argtypes = [vp, u32, ctypes.POINTER(u32)]
        H.solh_scene_material_textures.restype = vp
        H.solh_scene_material_textures.argtypes = [vp]
        H.solh_decode_png.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(u32), ctypes.POINTER(u32), vp, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        H.solh_mesh_info.argtypes = [vp, u32, ctypes.POINTER(_MeshInfo)]
        H.solh_mesh_sections.argtypes = [vp, u32, ctypes.POINTER(N.Section)]
        H.solh_mesh_extra_instance_count.restype = u32
        H.solh_mesh_extra_instance_count.argtypes = [vp, u32]
        H.solh_mesh_extra_instance_transform.argtypes = [vp, u32, u32, vp]
        H.solh_camera_new.restype = vp
        H.solh_camera_new.argtypes = [f, f]
        H.solh_camera_from_scene.restype = vp
        H.solh_camera_from_scene.argtypes = [vp]
        H.solh_camera_from_view.restype = vp
        H.solh_camera_from_view.argtypes = [vp, f, f, f]
        H.solh_camera_free.argtypes = [vp]
        H.solh_camera_look_at.argtypes = [vp, vp, vp, vp]
        H.solh_camera_set_window_size.argtypes = [vp, f, f]
        H.solh_camera_set_vfov.argtypes = [vp, f]
        H.solh_camera_matrices.argtypes = [vp, vp, vp]
        H.solh_scene_uniforms.argtypes = [vp, u32, u32, u32, ctypes.POINTER(N.SceneUniforms)]
        H.solh_mat4_inverse.argtypes = [vp, vp]
        H.solh_mat4_mul.argtypes = [vp, vp, vp]
        H.solh_mat4_from_scale_rotation_x.argtypes = [f, f, vp]
        _H = H
    return _H


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class PrimitiveSection:
    """src/scene/mod.rs:37-44"""

    def __init__(self, index, first_vertex, n_vertices, first_index, n_indices, material_index):
        self.index = index
        self.first_vertex, self.n_vertices = first_vertex, n_vertices
        self.first_index, self.n_indices = first_index, n_indices  # n_indices == 0: not indexed
        self.material_index = material_index  # None: no material


class Mesh:
    """src/scene/mesh.rs:53-61 with the buffers as host arrays: vertices float32 [n, 16] (ModelVertex),
    indices uint32 [m] (section-relative), transform float32 [16] column-major."""

    def __init__(self, name, vertices, indices, transform, primitive_sections, extra_instance_transforms=()):
        self.name = name
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 16)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self.transform = np.ascontiguousarray(transform, dtype=np.float32).reshape(16)
        self.primitive_sections = list(primitive_sections)
        # beyond the reference: global transforms of the other nodes referencing this mesh; the reference keeps
        # only the first node's (`transform`).  ray.SceneDescription.from_scene(..., instancing=True) instantiates them.
        self.extra_instance_transforms = [np.ascontiguousarray(t, dtype=np.float32).reshape(16) for t in extra_instance_transforms]


class Camera:
    """src/scene/camera.rs:34-127 (matrices only)."""

    def __init__(self, window_size=None, _handle=None):
        self._H = hostlib()
        self._h = _handle if _handle is not None else self._H.solh_camera_new(float(window_size[0]), float(window_size[1]))

    @classmethod
    def from_view(cls, view, yfov, z_near, z_far):
        v = np.ascontiguousarray(view, dtype=np.float32).reshape(16)
        return cls(_handle=hostlib().solh_camera_from_view(_fp(v), float(yfov), float(z_near), float(z_far)))

    def look_at(self, eye, center, up):
        e, c, u = (np.ascontiguousarray(x, dtype=np.float32) for x in (eye, center, up))
        self._H.solh_camera_look_at(self._h, _fp(e), _fp(c), _fp(u))

    def set_window_size(self, size):
        self._H.solh_camera_set_window_size(self._h, float(size[0]), float(size[1]))

    def set_vfov(self, vfov):
        self._H.solh_camera_set_vfov(self._h, float(vfov))

    def _mats(self):
        v = np.zeros(16, dtype=np.float32)
        p = np.zeros(16, dtype=np.float32)
        self._H.solh_camera_matrices(self._h, _fp(v), _fp(p))
        return v, p

    def view_matrix(self):
        return self._mats()[0]

    def perspective_matrix(self):
        return self._mats()[1]

    def __del__(self):
        try:
            if self._h:
                self._H.solh_camera_free(self._h)
                self._h = None
        except Exception:
            pass


def scene_uniforms(camera, width, height, frame):
    """SceneUniforms::from(camera, uvec3(w, h, elapsed_ticks)) — examples/5-pathtrace.rs:19-31,297-301."""
    u = N.SceneUniforms()
    hostlib().solh_scene_uniforms(camera._h, int(width), int(height), int(frame) & 0xFFFFFFFF, ctypes.byref(u))
    return u


class Texture:
    """Base-colour texture of a glTF material (the reference's load_scene reads no images): rgba8 [h, w, 4], rows
    top first, with the sampler's wrap modes (glTF codes, 10497 = REPEAT)."""

    def __init__(self, rgba8, wrap_s=10497, wrap_t=10497):
        self.rgba8 = np.ascontiguousarray(rgba8, dtype=np.uint8)
        assert self.rgba8.ndim == 3 and self.rgba8.shape[2] == 4
        self.wrap_s, self.wrap_t = int(wrap_s), int(wrap_t)


class Scene:
    """src/scene/mod.rs:99-104 (+ textures / material_textures, beyond the reference: bound with
    SceneDescription.set_textures or from_scene(..., textures=True); nothing samples them otherwise)"""

    def __init__(self, meshes, materials, camera=None, textures=(), material_textures=None):
        self.meshes = list(meshes)
        self.materials = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 12)  # MaterialInfo rows
        self.camera = camera
        self.textures = list(textures)
        self.material_textures = [None] * len(self.materials) if material_textures is None else list(material_textures)


def decode_png(data):
    """sol::image::decode_png (csrc/host/png.cpp) -> uint8 [h, w, 4]"""
    H = hostlib()
    buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
    w, h, err = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.create_string_buffer(256)
    if H.solh_decode_png(buf, len(data), ctypes.byref(w), ctypes.byref(h), None, 0, err, len(err)) != 0:
        raise N.SolbError(-1, err.value.decode("utf-8", "replace"))
    out = np.zeros((h.value, w.value, 4), dtype=np.uint8)
    H.solh_decode_png(buf, len(data), ctypes.byref(w), ctypes.byref(h), out.ctypes.data_as(ctypes.c_void_p), out.nbytes, err, len(err))
    return out


def load_scene(context, filepath):
    """scene::load_scene(context, path) — src/scene/mod.rs:138.  Raises SolbError where the reference panics."""
    H = hostlib()
    err = ctypes.create_string_buffer(512)
    h = H.solh_load_scene(os.fsencode(str(filepath)), err, len(err))
    if not h:
        raise N.SolbError(-1, err.value.decode("utf-8", "replace"))
    try:
        meshes = []
        for i in range(H.solh_scene_mesh_count(h)):
            info = _MeshInfo()
            H.solh_mesh_info(h, i, ctypes.byref(info))
            secs = (N.Section * max(info.n_sections, 1))()
            H.solh_mesh_sections(h, i, secs)
            verts = np.ctypeslib.as_array(ctypes.cast(info.vertices, ctypes.POINTER(ctypes.c_float)),
                                          shape=(info.n_vertices, 16)).copy() if info.n_vertices else np.zeros((0, 16), np.float32)
            inds = np.ctypeslib.as_array(ctypes.cast(info.indices, ctypes.POINTER(ctypes.c_uint32)),
                                         shape=(info.n_indices,)).copy() if info.n_indices else np.zeros((0,), np.uint32)
            sections = [PrimitiveSection(k, s.first_vertex, s.n_vertices, s.first_index, s.n_indices,
                                         None if s.material_index == 0xFFFFFFFF else int(s.material_index))
                        for k, s in enumerate(secs[: info.n_sections])]
            extra = []
            for k in range(H.solh_mesh_extra_instance_count(h, i)):
                t = np.zeros(16, dtype=np.float32)
                H.solh_mesh_extra_instance_transform(h, i, k, t.ctypes.data_as(ctypes.c_void_p))
                extra.append(t)
            meshes.append(Mesh(info.name.decode("utf-8", "replace"), verts, inds, np.array(info.transform[:], dtype=np.float32), sections,
                               extra))
        nm = H.solh_scene_material_count(h)
        mats = np.ctypeslib.as_array(ctypes.cast(H.solh_scene_materials(h), ctypes.POINTER(ctypes.c_float)),
                                     shape=(nm, 12)).copy() if nm else np.zeros((0, 12), np.float32)
        cam = Camera(_handle=H.solh_camera_from_scene(h)) if H.solh_scene_has_camera(h) else None
        textures = []
        for i in range(H.solh_scene_texture_count(h)):
            info = (ctypes.c_uint32 * 4)()
            px = H.solh_scene_texture(h, i, info)
            rgba = np.ctypeslib.as_array(ctypes.cast(px, ctypes.POINTER(ctypes.c_uint8)), shape=(info[1], info[0], 4)).copy()
            textures.append(Texture(rgba, info[2], info[3]))
        mt = np.ctypeslib.as_array(ctypes.cast(H.solh_scene_material_textures(h), ctypes.POINTER(ctypes.c_uint32)), shape=(nm,)).copy() if nm else []
        return Scene(meshes, mats, cam, textures, [None if int(t) == N.NO_TEXTURE else int(t) for t in mt])
    finally:
        H.solh_scene_free(h)
