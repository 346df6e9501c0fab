"""CPU oracle for the sol-rs ray-tracing hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (sol_rs_b200 / libsolb.so) never does.  PARITY UNPINNED — see
the header of oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import camera, gltf_flatten  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MISS = 0xFFFFFFFF
FLAG_EDGE, FLAG_TIE, FLAG_NEAR = 1, 2, 4


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class OrcInstance(ctypes.Structure):
    _fields_ = [
        ("first_vertex", ctypes.c_uint32), ("n_vertices", ctypes.c_uint32),
        ("first_index", ctypes.c_uint32), ("n_indices", ctypes.c_uint32),
        ("transform", ctypes.c_float * 16), ("transform_it", ctypes.c_float * 16),
        ("material", ctypes.c_float * 12),
    ]


class OrcStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("rays", "hits", "paths", "capped", "emissive")]


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_tea.restype = ctypes.c_uint32
        L.orc_tea.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.orc_next_word.restype = ctypes.c_uint32
        L.orc_next_word.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        L.orc_next_rand.restype = ctypes.c_float
        L.orc_next_rand.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        L.orc_fresnel_dielectric.restype = ctypes.c_float
        L.orc_fresnel_dielectric.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float]
        L.orc_sample_ggx.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p]
        L.orc_sample_cosine.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_scene_create.restype = ctypes.c_void_p
        L.orc_scene_create.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                       ctypes.c_void_p, ctypes.c_uint32]
        L.orc_scene_destroy.argtypes = [ctypes.c_void_p]
        L.orc_scene_tri_count.restype = ctypes.c_uint32
        L.orc_scene_tri_count.argtypes = [ctypes.c_void_p]
        L.orc_scene_bounds.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_trace_rays.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        L.orc_debug.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        L.orc_pathtrace_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                          ctypes.c_int32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]
        L.orc_ao_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int32,
                                   ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        _LIB = L
    return _LIB


def tea(a, b):
    return lib().orc_tea(a & 0xFFFFFFFF, b & 0xFFFFFFFF)


def rand_stream(seed, n):
    """n successive (word, float) draws of nextRand from state `seed`."""
    st = ctypes.c_uint32(seed)
    out = []
    for _ in range(n):
        st2 = ctypes.c_uint32(st.value)
        w = lib().orc_next_word(ctypes.byref(st2))
        f = lib().orc_next_rand(ctypes.byref(st))
        out.append((w, f))
    return out


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


DEFAULT_EPS_B = 1e-4   # barycentric distance from an edge below which a primary hit is "edge"
DEFAULT_EPS_T = 1e-5   # relative t gap below which two hits are a "tie"


class Scene:
    """Oracle scene built from a gltf_flatten.FlatScene (reference-layout buffers)."""

    def __init__(self, flat, transforms=None):
        self.flat = flat
        n = len(flat.instances)
        arr = (OrcInstance * max(n, 1))()
        for i, inst in enumerate(flat.instances):
            t = np.asarray(transforms[i] if transforms is not None else inst["transform"], dtype=np.float32)
            tit = gltf_flatten.mat4_inverse(t).T.copy()  # inverse().transpose(): src/ray/mod.rs:116
            arr[i].first_vertex = inst["first_vertex"]
            arr[i].n_vertices = inst["n_vertices"]
            arr[i].first_index = inst["first_index"]
            arr[i].n_indices = inst["n_indices"]
            arr[i].transform[:] = t.reshape(16).tolist()
            arr[i].transform_it[:] = tit.reshape(16).tolist()
            arr[i].material[:] = flat.materials[inst["material"]].tolist()
        self._inst = arr
        self.vertices = np.ascontiguousarray(flat.vertices, dtype=np.float32)
        self.indices = np.ascontiguousarray(flat.indices, dtype=np.uint32)
        self.handle = lib().orc_scene_create(n, arr, _ptr(self.vertices), self.vertices.shape[0],
                                             _ptr(self.indices), self.indices.shape[0])

    def set_textures(self, textures, material_textures):
        """EXTENSION shared with the product, not reference behaviour: textures = [(rgba8 [h, w, 4] rows top first,
        wrap_s, wrap_t)], material_textures = per material an index or None; instance i samples the texture of its material."""
        n = len(textures)
        imgs = [np.ascontiguousarray(t[0], dtype=np.uint8) for t in textures]
        ptrs = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in imgs])
        w = np.array([a.shape[1] for a in imgs], dtype=np.uint32)
        h = np.array([a.shape[0] for a in imgs], dtype=np.uint32)
        ws = np.array([t[1] for t in textures], dtype=np.uint32)
        wt = np.array([t[2] for t in textures], dtype=np.uint32)
        it = np.array([0xFFFFFFFF if material_textures[inst["material"]] is None else material_textures[inst["material"]]
                       for inst in self.flat.instances], dtype=np.uint32)
        L = lib()
        L.orc_scene_set_textures.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.orc_scene_set_textures(self.handle, n, ptrs, _ptr(w), _ptr(h), _ptr(ws), _ptr(wt), 1, _ptr(it))
        self.instance_textures = it

    def sample_texture(self, t, u, v):
        out = np.zeros(3, dtype=np.float32)
        L = lib()
        L.orc_sample_texture.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
        L.orc_sample_texture(self.handle, int(t), float(u), float(v), _ptr(out))
        return out

    def __del__(self):
        if getattr(self, "handle", None):
            lib().orc_scene_destroy(self.handle)
            self.handle = None

    @property
    def tri_count(self):
        return lib().orc_scene_tri_count(self.handle)

    def bounds(self):
        lo = np.zeros(3, dtype=np.float64)
        hi = np.zeros(3, dtype=np.float64)
        lib().orc_scene_bounds(self.handle, _ptr(lo), _ptr(hi))
        return lo, hi

    def trace_rays(self, rays, classify=False, eps_b=DEFAULT_EPS_B, eps_t=DEFAULT_EPS_T):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        hits = np.zeros((n, 4), dtype=np.uint32)
        t = np.zeros(n, dtype=np.float32)
        flags = np.zeros(n, dtype=np.uint8) if classify else None
        lib().orc_trace_rays(self.handle, _ptr(rays), n, _ptr(hits), _ptr(t), _ptr(flags), eps_b, eps_t)
        return hits, t, flags

    def debug(self, uniforms, w, h, eps_b=DEFAULT_EPS_B, eps_t=DEFAULT_EPS_T):
        rgba = np.zeros((h, w, 4), dtype=np.uint8)
        ids = np.zeros((h, w, 2), dtype=np.uint32)
        bary_t = np.zeros((h, w, 3), dtype=np.float32)
        flags = np.zeros((h, w), dtype=np.uint8)
        ub = ctypes.create_string_buffer(uniforms, 400)
        lib().orc_debug(self.handle, ub, w, h, _ptr(rgba), _ptr(ids), _ptr(bary_t), _ptr(flags), eps_b, eps_t)
        return rgba, ids, bary_t, flags

    def pathtrace_frame(self, uniforms, w, h, accum, accum_start_frame=0, enable_sky=False, spp=8, max_bounces=32,
                        stats=None):
        assert accum.dtype == np.float32 and accum.shape == (h, w, 4) and accum.flags.c_contiguous
        render = np.zeros((h, w, 4), dtype=np.uint8)
        ub = ctypes.create_string_buffer(uniforms, 400)
        st = stats if stats is not None else OrcStats()
        lib().orc_pathtrace_frame(self.handle, ub, w, h, accum_start_frame, int(enable_sky), spp, max_bounces,
                                  _ptr(accum), _ptr(render), ctypes.byref(st))
        return render, st

    def ao_frame(self, uniforms, w, h, image, blue, accum_start_frame=0, stats=None):
        assert image.dtype == np.float32 and image.shape == (h, w, 4) and image.flags.c_contiguous
        blue = np.ascontiguousarray(blue, dtype=np.uint8)
        ub = ctypes.create_string_buffer(uniforms, 400)
        st = stats if stats is not None else OrcStats()
        lib().orc_ao_frame(self.handle, ub, w, h, accum_start_frame, _ptr(blue), blue.shape[1], blue.shape[0],
                           _ptr(image), ctypes.byref(st))
        return st
