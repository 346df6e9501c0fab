"""ctypes driver for tests/emu/emu.cpp — the TEST-ONLY host emulation of libsolb's per-thread device
logic (builder, traversal, shading).  Not a product path; see the header of emu.cpp."""
import ctypes
import os
import subprocess

import numpy as np

import oracle
from oracle import gltf_flatten as gf

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None


class DeviceInstance(ctypes.Structure):
    _fields_ = [("first_vertex", ctypes.c_uint32), ("first_index", ctypes.c_uint32), ("n_indices", ctypes.c_uint32),
                ("material", ctypes.c_uint32), ("blas", ctypes.c_uint32), ("shade_first_tri", ctypes.c_uint32),
                ("transform", ctypes.c_float * 16), ("transform_it", ctypes.c_float * 16), ("mat", ctypes.c_float * 12)]


assert ctypes.sizeof(DeviceInstance) == 200  # solb_internal.h


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "emu", "libemu.so")
        srcs = [os.path.join(_HERE, "emu", "emu.cpp")] + [os.path.join(_ROOT, "sol_rs_b200", "csrc", f)
                                                          for f in ("bvh.cuh", "build.cuh", "shade.cuh", "common.cuh", "solb_internal.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-I",
                                   os.path.join(_ROOT, "sol_rs_b200", "csrc"), "-I", "/usr/local/cuda/include", "-o", so, srcs[0]])
        L = ctypes.CDLL(so)
        L.emu_scene_create.restype = ctypes.c_void_p
        L.emu_scene_create.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32]
        L.emu_scene_destroy.argtypes = [ctypes.c_void_p]
        L.emu_build.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.emu_build_two_level.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.emu_set_dp_collapse.argtypes = [ctypes.c_int]
        L.emu_set_transform.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        for f in ("emu_node_count", "emu_depth"):
            getattr(L, f).restype = ctypes.c_uint32
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.emu_max_stack.argtypes = [ctypes.c_void_p]
        L.emu_sah.restype = ctypes.c_float
        L.emu_sah.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.emu_read_nodes.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.emu_read_tris.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.emu_trace_rays.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.emu_debug.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        L.emu_pathtrace_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.emu_mask_selftest.restype = ctypes.c_uint32
        L.emu_mask_selftest.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.emu_tea.restype = ctypes.c_uint32
        L.emu_tea.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.emu_next_rand.restype = ctypes.c_float
        L.emu_next_rand.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class EmuScene:
    def __init__(self, flat, treelet_passes=2, gamma=7, two_level=False, extra_instances=(), dp_collapse=True):
        """extra_instances: (source_instance, 4x4 transform, material_index) tuples = solb_scene_add_instance."""
        insts = [dict(inst, blas=i) for i, inst in enumerate(flat.instances)]
        for src, t, mat in extra_instances:
            insts.append(dict(insts[src], transform=np.asarray(t, dtype=np.float32), material=mat))
        n = len(insts)
        arr = (DeviceInstance * max(n, 1))()
        for i, inst in enumerate(insts):
            t = np.asarray(inst["transform"], dtype=np.float32)
            tit = gf.mat4_inverse(t).T.copy()
            arr[i].first_vertex = inst["first_vertex"]
            arr[i].first_index = inst["first_index"]
            arr[i].n_indices = inst["n_indices"]
            arr[i].material = inst["material"]
            arr[i].blas = inst["blas"]
            arr[i].transform[:] = t.reshape(16).tolist()
            arr[i].transform_it[:] = tit.reshape(16).tolist()
            arr[i].mat[:] = flat.materials[inst["material"]].tolist()
        v = np.ascontiguousarray(flat.vertices, dtype=np.float32)
        idx = np.ascontiguousarray(flat.indices, dtype=np.uint32)
        self.h = lib().emu_scene_create(n, arr, _p(v), v.shape[0], _p(idx), idx.shape[0])
        self.passes, self.gamma, self.two_level, self.dp_collapse = treelet_passes, gamma, two_level, dp_collapse
        self.build()

    def build(self):
        lib().emu_set_dp_collapse(int(self.dp_collapse))
        rc = (lib().emu_build_two_level if self.two_level else lib().emu_build)(self.h, self.passes, self.gamma)
        assert rc == 0, "emu_build: triangle / instance count mismatch"

    def set_textures(self, textures, material_textures, flat):
        """textures = [(rgba8 [h, w, 4] rows top first, wrap_s, wrap_t)], material_textures = per material an index or None"""
        n = len(textures)
        imgs = [np.ascontiguousarray(t[0], dtype=np.uint8) for t in textures]
        ptrs = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in imgs])
        w = np.array([a.shape[1] for a in imgs], dtype=np.uint32)
        h = np.array([a.shape[0] for a in imgs], dtype=np.uint32)
        ws = np.array([t[1] for t in textures], dtype=np.uint32)
        wt = np.array([t[2] for t in textures], dtype=np.uint32)
        it = np.array([0xFFFFFFFF if material_textures[i["material"]] is None else material_textures[i["material"]] for i in flat.instances],
                      dtype=np.uint32)
        L = lib()
        L.emu_scene_set_textures.argtypes = [ctypes.c_void_p, ctypes.c_uint32] + [ctypes.c_void_p] * 6
        L.emu_scene_set_textures(self.h, n, ptrs, _p(w), _p(h), _p(ws), _p(wt), _p(it))

    def set_transform(self, index, transform):
        t = np.ascontiguousarray(transform, dtype=np.float32).reshape(4, 4)
        tit = np.ascontiguousarray(gf.mat4_inverse(t).T, dtype=np.float32)
        lib().emu_set_transform(self.h, index, _p(t), _p(tit))

    def __del__(self):
        if getattr(self, "h", None):
            lib().emu_scene_destroy(self.h)
            self.h = None

    def nodes(self):
        out = np.zeros((lib().emu_node_count(self.h), 20), dtype=np.uint32)
        lib().emu_read_nodes(self.h, _p(out))
        return out

    def tris(self, n):
        out = np.zeros((n, 12), dtype=np.float32)
        lib().emu_read_tris(self.h, _p(out))
        return out

    def info(self):
        L = lib()
        return dict(nodes=L.emu_node_count(self.h), depth=L.emu_depth(self.h), sah_lbvh=L.emu_sah(self.h, 0),
                    sah=L.emu_sah(self.h, 1), max_stack=L.emu_max_stack(self.h))

    def trace_rays(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        hits = np.zeros((n, 4), dtype=np.uint32)
        t = np.zeros(n, dtype=np.float32)
        ctr = np.zeros(2, dtype=np.uint64)
        lib().emu_trace_rays(self.h, _p(rays), n, _p(hits), _p(t), _p(ctr))
        return hits, t, ctr

    def debug(self, uniforms, w, h):
        u = np.frombuffer(uniforms, dtype=np.float32).copy()
        render = np.zeros((h, w), dtype=np.uint32)
        ids = np.zeros((h, w, 2), dtype=np.uint32)
        lib().emu_debug(self.h, _p(u), w, h, _p(render), _p(ids))
        return render.view(np.uint8).reshape(h, w, 4), ids

    def pathtrace_frame(self, uniforms, w, h, accum, accum_start=0, enable_sky=False, spp=8, max_bounces=32, accum_mode=0):
        u = np.frombuffer(uniforms, dtype=np.float32).copy()
        render = np.zeros((h, w), dtype=np.uint32)
        stats = np.zeros(4, dtype=np.uint64)  # rays, hits, wide nodes visited, triangles tested
        lib().emu_pathtrace_frame(self.h, _p(u), w, h, accum_start, int(enable_sky), spp, max_bounces, accum_mode, _p(accum),
                                  _p(render), _p(stats))
        return render.view(np.uint8).reshape(h, w, 4), stats
