"""ctypes binding of libsolb.so (include/solb.h).  The library is built in-tree by
`__graft_entry__.build()` / `make -C sol_rs_b200/csrc`; there is no fallback: if it is missing or a call
fails, SolbError is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SOLB_LIB_PATH: load an experimental build of the same library instead (A/B runs of compile-time variants)
LIB_PATH = os.environ.get("SOLB_LIB_PATH") or os.path.join(_HERE, "libsolb.so")

SOLB_OK = 0
FORMAT_RGBA32F, FORMAT_RGBA8, FORMAT_RG32UI = 0, 1, 2
SCHEDULE_WAVEFRONT, SCHEDULE_MEGAKERNEL, SCHEDULE_AUTO, SCHEDULE_WARPFRONT = 0, 1, 2, 3
ACCUM_MIX, ACCUM_SUM = 0, 1
ACCEL_FLAT, ACCEL_TWO_LEVEL = 0, 1
MISS = 0xFFFFFFFF


class SolbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsolb error %d: %s" % (code, msg))
        self.code = code


class ModelVertex(ctypes.Structure):
    _fields_ = [("pos", ctypes.c_float * 4), ("color", ctypes.c_float * 4), ("normal", ctypes.c_float * 4), ("uv", ctypes.c_float * 4)]


class MaterialInfo(ctypes.Structure):
    _fields_ = [("base_color", ctypes.c_float * 4), ("emissive", ctypes.c_float * 3), ("padding0", ctypes.c_float),
                ("metallic", ctypes.c_float), ("roughness", ctypes.c_float), ("padding1", ctypes.c_float), ("padding2", ctypes.c_float)]


class SceneInstance(ctypes.Structure):
    _fields_ = [("id", ctypes.c_uint32), ("texture_offset", ctypes.c_uint32), ("padding", ctypes.c_float * 2),
                ("transform", ctypes.c_float * 16), ("transform_it", ctypes.c_float * 16)]


class SceneUniforms(ctypes.Structure):
    _fields_ = [("model", ctypes.c_float * 16), ("view", ctypes.c_float * 16), ("view_inverse", ctypes.c_float * 16),
                ("projection", ctypes.c_float * 16), ("projection_inverse", ctypes.c_float * 16),
                ("model_view_projection", ctypes.c_float * 16), ("frame", ctypes.c_uint32 * 3), ("_pad", ctypes.c_uint32)]


class Section(ctypes.Structure):
    _fields_ = [("first_vertex", ctypes.c_uint32), ("n_vertices", ctypes.c_uint32), ("first_index", ctypes.c_uint32),
                ("n_indices", ctypes.c_uint32), ("material_index", ctypes.c_uint32)]


class MeshDesc(ctypes.Structure):
    _fields_ = [("vertices", ctypes.c_void_p), ("n_vertices", ctypes.c_uint32), ("indices", ctypes.c_void_p),
                ("n_indices", ctypes.c_uint32), ("sections", ctypes.POINTER(Section)), ("n_sections", ctypes.c_uint32),
                ("transform", ctypes.c_float * 16)]


class TextureDesc(ctypes.Structure):
    _fields_ = [("rgba8", ctypes.c_void_p), ("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("wrap_s", ctypes.c_uint32),
                ("wrap_t", ctypes.c_uint32), ("srgb", ctypes.c_uint32), ("_pad", ctypes.c_uint32)]


NO_TEXTURE = 0xFFFFFFFF


class TraceParams(ctypes.Structure):
    _fields_ = [("accum_start_frame", ctypes.c_int32), ("enable_sky", ctypes.c_uint32), ("samples_per_frame", ctypes.c_uint32),
                ("max_bounces", ctypes.c_uint32), ("schedule", ctypes.c_uint32), ("accum_mode", ctypes.c_uint32),
                ("collect_stats", ctypes.c_uint32), ("tile_row_begin", ctypes.c_uint32), ("tile_row_count", ctypes.c_uint32),
                ("tile_row_stride", ctypes.c_uint32)]


class Stats(ctypes.Structure):
    _fields_ = [("rays", ctypes.c_uint64), ("hits", ctypes.c_uint64), ("paths", ctypes.c_uint64), ("nodes_visited", ctypes.c_uint64),
                ("tris_tested", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64), ("last_build_ms", ctypes.c_float),
                ("last_trace_ms", ctypes.c_float), ("trace_kernel_ms_total", ctypes.c_float), ("trace_kernel_launches", ctypes.c_uint32)]


class AccelInfo(ctypes.Structure):
    _fields_ = [("n_instances", ctypes.c_uint32), ("n_triangles", ctypes.c_uint32), ("n_wide_nodes", ctypes.c_uint32),
                ("wide_depth", ctypes.c_uint32), ("n_binary_nodes", ctypes.c_uint32), ("sah_cost_binary", ctypes.c_float),
                ("sah_cost_lbvh", ctypes.c_float), ("scene_lo", ctypes.c_float * 3), ("scene_hi", ctypes.c_float * 3),
                ("mode", ctypes.c_uint32), ("n_blas", ctypes.c_uint32), ("n_tlas_nodes", ctypes.c_uint32), ("tlas_depth", ctypes.c_uint32)]


assert ctypes.sizeof(ModelVertex) == 64 and ctypes.sizeof(MaterialInfo) == 48
assert ctypes.sizeof(SceneInstance) == 144 and ctypes.sizeof(SceneUniforms) == 400

# every symbol include/solb.h declares: (restype, argtypes)
_vp, _u32, _i = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
_pp = ctypes.POINTER(ctypes.c_void_p)
SYMBOLS = {
    "solb_ctx_create": (_i, [_i, _vp, _pp]),
    "solb_ctx_destroy": (_i, [_vp]),
    "solb_ctx_preload": (_i, [_vp]),
    "solb_synchronize": (_i, [_vp]),
    "solb_fence_create": (_i, [_vp, _pp]),
    "solb_fence_signal": (_i, [_vp]),
    "solb_fence_wait": (_i, [_vp]),
    "solb_fence_destroy": (_i, [_vp]),
    "solb_host_alloc": (_i, [_vp, ctypes.c_size_t, _pp]),
    "solb_host_free": (_i, [_vp, _vp]),
    "solb_ctx_trim": (_i, [_vp]),
    "solb_last_error": (ctypes.c_char_p, [_vp]),
    "solb_version": (_u32, []),
    "solb_stats_get": (_i, [_vp, ctypes.POINTER(Stats)]),
    "solb_stats_reset": (_i, [_vp]),
    "solb_set_timing": (_i, [_vp, _i]),
    "solb_scene_create": (_i, [_vp, ctypes.POINTER(MeshDesc), _u32, ctypes.POINTER(MaterialInfo), _u32, _pp]),
    "solb_scene_destroy": (_i, [_vp]),
    "solb_accel_build": (_i, [_vp]),
    "solb_instance_set_transform": (_i, [_vp, _u32, ctypes.POINTER(ctypes.c_float)]),
    "solb_scene_update": (_i, [_vp]),
    "solb_tlas_regenerate": (_i, [_vp]),
    "solb_scene_set_textures": (_i, [_vp, ctypes.POINTER(TextureDesc), _u32, ctypes.POINTER(_u32), _u32]),
    "solb_scene_add_instance": (_i, [_vp, _u32, ctypes.POINTER(ctypes.c_float), _u32, ctypes.POINTER(_u32)]),
    "solb_scene_set_accel_mode": (_i, [_vp, _u32]),
    "solb_scene_instance_count": (_i, [_vp, ctypes.POINTER(_u32)]),
    "solb_scene_get_instances": (_i, [_vp, ctypes.POINTER(SceneInstance), _u32]),
    "solb_scene_instance_triangles": (_i, [_vp, ctypes.POINTER(_u32), _u32]),
    "solb_accel_info": (_i, [_vp, ctypes.POINTER(AccelInfo)]),
    "solb_accel_read_nodes": (_i, [_vp, _vp, ctypes.c_size_t]),
    "solb_accel_read_triangles": (_i, [_vp, _vp, ctypes.c_size_t]),
    "solb_target_create": (_i, [_vp, _u32, _u32, _u32, _pp]),
    "solb_target_destroy": (_i, [_vp]),
    "solb_target_clear": (_i, [_vp]),
    "solb_target_readback": (_i, [_vp, _vp, ctypes.c_size_t]),
    "solb_target_readback_async": (_i, [_vp, _vp, ctypes.c_size_t]),
    "solb_target_upload": (_i, [_vp, _vp, ctypes.c_size_t]),
    "solb_target_device_ptr": (_i, [_vp, _pp]),
    "solb_target_info": (_i, [_vp, ctypes.POINTER(_u32), ctypes.POINTER(_u32), ctypes.POINTER(_u32)]),
    "solb_trace_params_default": (None, [ctypes.POINTER(TraceParams), _i]),
    "solb_trace_pathtrace": (_i, [_vp, ctypes.POINTER(SceneUniforms), ctypes.POINTER(TraceParams), _vp, _vp]),
    "solb_set_blue_noise": (_i, [_vp, _vp, _u32, _u32]),
    "solb_trace_ao": (_i, [_vp, ctypes.POINTER(SceneUniforms), ctypes.POINTER(TraceParams), _vp]),
    "solb_trace_debug": (_i, [_vp, ctypes.POINTER(SceneUniforms), _vp, _vp, _vp]),
    "solb_trace_rays": (_i, [_vp, _vp, _u32, _vp, _vp]),
    "solb_resolve_sum": (_i, [_vp, _vp, _vp, _vp]),
    "solb_test_sort_pairs": (_i, [_vp, _vp, _vp, _u32, _i]),
    "solb_comm_unique_id": (_i, [_vp]),
    "solb_comm_init": (_i, [_vp, _vp, _i, _i]),
    "solb_comm_info": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "solb_comm_destroy": (_i, [_vp]),
    "solb_reduce_accum": (_i, [_vp, _vp, _i, _vp, _vp]),
    "solb_allgather_rows": (_i, [_vp, _vp, _u32]),
}
COMM_ID_BYTES = 128

_LIB = None


def lib():
    """Load libsolb.so (raises if it has not been built: there is no fallback path)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SolbError(-2, "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(libsolb has no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc, ctx=None):
    if rc != SOLB_OK:
        msg = lib().solb_last_error(ctx)
        raise SolbError(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc
