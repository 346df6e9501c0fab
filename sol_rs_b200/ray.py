"""sol::ray mirror: SceneDescription, PipelineInfo, Pipeline, ShaderBindingTableInfo, ShaderBindingTable.

Names, call order and argument meaning follow /root/reference src/ray/{mod,pipeline,sbt}.rs and the
setup()/render() callbacks of examples/{3-ray-debug,4-ray-ao,5-pathtrace}.rs; every call lands in the
C ABI of libsolb.so (include/solb.h).  No CPU path exists."""
import ctypes
import os

import numpy as np

from . import _native as N

RAYGEN_KHR, MISS_KHR, CLOSEST_HIT_KHR = "raygen", "miss", "closest_hit"
PATHTRACE, AO, DEBUG = "pathtrace", "ao", "debug"


class SceneDescription:
    """src/ray/mod.rs:38-206.  One BLAS + one instance per primitive section, in mesh x section order."""

    def __init__(self, context, handle):
        self.context = context
        self._h = handle
        self._lib = N.lib()

    @classmethod
    def from_scene(cls, context, scene, accel_mode=N.ACCEL_FLAT, instancing=False, textures=False):
        """src/ray/mod.rs:50-57.  instancing=True (beyond the reference): every further glTF node that references
        a mesh becomes one more instance of that mesh's BLASes (instance ids continue the running count).  textures=True
        (beyond the reference): bind the scene's base-colour textures (set_textures)."""
        sd = cls.from_meshes(context, scene.meshes, [m.transform for m in scene.meshes], scene.materials, accel_mode)
        if textures and getattr(scene, "textures", None):
            sd.set_textures(scene.textures, scene.material_textures)
        if instancing:
            first, added = 0, False
            for m in scene.meshes:
                for t in getattr(m, "extra_instance_transforms", ()):
                    for k, ps in enumerate(m.primitive_sections):
                        sd.add_instance(first + k, t, ps.material_index)
                        added = True
                first += len(m.primitive_sections)
            if added:
                sd.accel_build()
        return sd

    @classmethod
    def from_meshes(cls, context, meshes, mesh_transforms, materials, accel_mode=N.ACCEL_FLAT):
        """accel_mode: N.ACCEL_FLAT (default, transforms baked into one hierarchy) or N.ACCEL_TWO_LEVEL (TLAS over shared
        object-space BLASes, the reference's own split: src/ray/acceleration.rs)."""
        assert len(meshes) == len(mesh_transforms)
        L = N.lib()
        descs = (N.MeshDesc * max(len(meshes), 1))()
        keep = []
        for i, m in enumerate(meshes):
            secs = (N.Section * max(len(m.primitive_sections), 1))()
            for k, ps in enumerate(m.primitive_sections):
                if ps.material_index is None:  # the reference unwrap()s (src/scene/mod.rs:65)
                    raise N.SolbError(-1, "primitive without material (Option::unwrap on None)")
                secs[k] = N.Section(ps.first_vertex, ps.n_vertices, ps.first_index, ps.n_indices, ps.material_index)
            v = np.ascontiguousarray(m.vertices, dtype=np.float32)
            idx = np.ascontiguousarray(m.indices, dtype=np.uint32)
            keep += [secs, v, idx]
            d = descs[i]
            d.vertices = v.ctypes.data if v.size else None
            d.n_vertices = v.shape[0]
            d.indices = idx.ctypes.data if idx.size else None
            d.n_indices = idx.shape[0]
            d.sections = secs
            d.n_sections = len(m.primitive_sections)
            d.transform[:] = np.asarray(mesh_transforms[i], dtype=np.float32).reshape(16).tolist()
        mats = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 12) if materials is not None else np.zeros((0, 12), np.float32)
        h = ctypes.c_void_p()
        N.check(L.solb_scene_create(context.handle, descs, len(meshes),
                                    ctypes.cast(mats.ctypes.data, ctypes.POINTER(N.MaterialInfo)) if mats.size else None,
                                    mats.shape[0], ctypes.byref(h)), context.handle)
        sd = cls(context, h)
        try:
            N.check(L.solb_scene_set_accel_mode(h, int(accel_mode)), context.handle)
            N.check(L.solb_accel_build(h), context.handle)  # BLAS::new x n + TLAS::new + end_single_time_cmd
        except Exception:
            sd.close()
            raise
        return sd

    def blas_transform(self, transform, index):
        t = np.ascontiguousarray(transform, dtype=np.float32).reshape(16)
        N.check(self._lib.solb_instance_set_transform(self._h, int(index), t.ctypes.data_as(ctypes.POINTER(ctypes.c_float))),
                self.context.handle)

    def add_instance(self, source_instance, transform, material_index):
        """One more instance of the BLAS `source_instance` uses (the reference's TODO at src/ray/mod.rs:122).
        Returns the new gl_InstanceID.  Call accel_build() / tlas_regenerate() afterwards."""
        t = np.ascontiguousarray(transform, dtype=np.float32).reshape(16)
        out = ctypes.c_uint32()
        N.check(self._lib.solb_scene_add_instance(self._h, int(source_instance), t.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                                  int(material_index), ctypes.byref(out)), self.context.handle)
        return out.value

    def set_textures(self, textures, material_textures):
        """solb_scene_set_textures: textures = scene.Texture list (rgba8 rows top first, sRGB-encoded colour), material_textures
        = per material a texture index or None.  Every instance's texture_offset (src/ray/mod.rs:20) becomes the texture of its
        material and 5-pathtrace multiplies the albedo by a bilinear sample at the hit's uv.  An empty list unbinds."""
        descs = (N.TextureDesc * max(len(textures), 1))()
        keep = []
        for i, t in enumerate(textures):
            a = np.ascontiguousarray(t.rgba8, dtype=np.uint8)
            keep.append(a)
            descs[i] = N.TextureDesc(a.ctypes.data, a.shape[1], a.shape[0], int(t.wrap_s), int(t.wrap_t), 1, 0)
        mt = np.array([N.NO_TEXTURE if t is None else int(t) for t in material_textures], dtype=np.uint32)
        N.check(self._lib.solb_scene_set_textures(self._h, descs, len(textures), mt.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(mt)),
                self.context.handle)

    def set_accel_mode(self, mode):
        N.check(self._lib.solb_scene_set_accel_mode(self._h, int(mode)), self.context.handle)

    def accel_build(self):
        N.check(self._lib.solb_accel_build(self._h), self.context.handle)

    def blas_transforms(self, transforms):
        for i, t in enumerate(transforms):
            self.blas_transform(t, i)

    def tlas_regenerate(self, cmd=None):
        N.check(self._lib.solb_tlas_regenerate(self._h), self.context.handle)

    def update(self):
        N.check(self._lib.solb_scene_update(self._h), self.context.handle)

    def blas_count(self):
        n = ctypes.c_uint32()
        N.check(self._lib.solb_scene_instance_count(self._h, ctypes.byref(n)), self.context.handle)
        return n.value

    def instance_triangles(self):
        """triangle count of every instance (= of the BLAS it places)"""
        return [int(n) for n in self._instance_tris()]

    def _instance_tris(self):
        n = self.blas_count()
        out = (ctypes.c_uint32 * max(n, 1))()
        N.check(self._lib.solb_scene_instance_triangles(self._h, out, n), self.context.handle)
        return out[:n]

    def instances(self):
        n = self.blas_count()
        arr = (N.SceneInstance * max(n, 1))()
        N.check(self._lib.solb_scene_get_instances(self._h, arr, n), self.context.handle)
        return list(arr[:n])

    def accel_info(self):
        info = N.AccelInfo()
        N.check(self._lib.solb_accel_info(self._h, ctypes.byref(info)), self.context.handle)
        return info

    def read_nodes(self):
        info = self.accel_info()
        out = np.zeros((info.n_wide_nodes, 20), dtype=np.uint32)
        N.check(self._lib.solb_accel_read_nodes(self._h, out.ctypes.data_as(ctypes.c_void_p), out.nbytes), self.context.handle)
        return out

    def read_triangles(self):
        info = self.accel_info()
        out = np.zeros((info.n_triangles, 12), dtype=np.float32)
        N.check(self._lib.solb_accel_read_triangles(self._h, out.ctypes.data_as(ctypes.c_void_p), out.nbytes), self.context.handle)
        return out

    def trace_rays(self, rays):
        """traceRayEXT for host rays [n, 8] = (o, tmin, d, tmax) -> hits uint32 [n, 4] (inst, prim, bits u, bits v), t [n]."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        hits = np.zeros((n, 4), dtype=np.uint32)
        t = np.zeros(n, dtype=np.float32)
        N.check(self._lib.solb_trace_rays(self._h, rays.ctypes.data_as(ctypes.c_void_p), n, hits.ctypes.data_as(ctypes.c_void_p),
                                          t.ctypes.data_as(ctypes.c_void_p)), self.context.handle)
        return hits, t

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            self._lib.solb_scene_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PipelineInfo:
    """src/ray/pipeline.rs:5-52"""

    def __init__(self):
        self.shaders, self.spec, self.spec_id, self._name = [], None, 0, ""

    def layout(self, _layout=None):
        return self

    def shader(self, path, stage):
        self.shaders.append((str(path), stage))
        return self

    def specialization(self, data, constant_id):
        self.spec, self.spec_id = list(data), int(constant_id)
        return self

    def name(self, name):
        self._name = name
        return self


class Pipeline:
    """ray::Pipeline::new (src/ray/pipeline.rs:61-122).  The CUDA kernels are compiled ahead of time, so this
    selects the kernel family from the raygen shader's file name and captures ENABLE_SKYLIGHT (constant 0)."""

    def __init__(self, context, info):
        self.context = context
        rgen = [p for p, s in info.shaders if s == RAYGEN_KHR]
        if len(rgen) != 1:
            raise N.SolbError(-1, "ray::Pipeline: exactly one raygen stage expected")
        base = os.path.basename(rgen[0])
        kinds = {"pathtrace.rgen": PATHTRACE, "ao.rgen": AO, "debug.rgen": DEBUG}
        if base not in kinds:
            raise N.SolbError(-4, "ray::Pipeline: no CUDA kernel family for raygen shader %r" % base)
        self.kind = kinds[base]
        self.enable_sky = bool(info.spec and info.spec_id == 0 and info.spec[0])
        # the reference compiles its GLSL stages here (src/ray/pipeline.rs:105-115); the counterpart is loading the kernels
        if context is not None:
            context.preload()


class ShaderBindingTableInfo:
    """src/ray/sbt.rs:13-55"""

    def __init__(self):
        self._raygen, self._miss, self._hit = [], [], []

    def raygen(self, i):
        self._raygen.append(i)
        return self

    def miss(self, i):
        self._miss.append(i)
        return self

    def hitgroup(self, i):
        self._hit.append(i)
        return self


class TraceBindings:
    """What the reference binds before cmd_trace_rays (examples/5-pathtrace.rs:297-356): uniforms (set 0),
    TLAS + images (set 1), push constant.  Optional overrides expose the shader literals BASELINE's configs vary."""

    def __init__(self, scene_description, uniforms, accum_target=None, render_target=None, ids_target=None,
                 accumulation_start_frame=0, samples_per_frame=None, max_bounces=None, schedule=N.SCHEDULE_AUTO,
                 accum_mode=N.ACCUM_MIX, collect_stats=False, tile_rows=None):
        self.scene_description, self.uniforms = scene_description, uniforms
        self.accum_target, self.render_target, self.ids_target = accum_target, render_target, ids_target
        self.accumulation_start_frame = accumulation_start_frame
        self.samples_per_frame, self.max_bounces = samples_per_frame, max_bounces
        self.schedule, self.accum_mode, self.collect_stats = schedule, accum_mode, collect_stats
        self.tile_rows = tile_rows  # (first_row, n_rows) of the full-size targets: tile split of one frame


class ShaderBindingTable:
    """src/ray/sbt.rs:58-181"""

    def __init__(self, context, pipeline, info):
        if len(info._raygen) != 1 or len(info._miss) != 1 or len(info._hit) != 1:
            raise N.SolbError(-4, "ShaderBindingTable: exactly one raygen / miss / hit group is supported")
        self.context, self.pipeline = context, pipeline
        self._lib = N.lib()

    def cmd_trace_rays(self, bindings, extent):
        """src/ray/sbt.rs:167-180.  extent = (width, height, 1) must match the bound targets."""
        b = bindings
        first = b.accum_target or b.render_target or b.ids_target
        if first is None:
            raise N.SolbError(-1, "cmd_trace_rays: no storage image bound")
        if tuple(extent) != (first.width, first.height, 1):
            raise N.SolbError(-1, "cmd_trace_rays: extent differs from the bound targets")
        p = N.TraceParams()
        self._lib.solb_trace_params_default(ctypes.byref(p), 1 if self.pipeline.kind == AO else 0)
        p.accum_start_frame = int(b.accumulation_start_frame)
        p.enable_sky = int(self.pipeline.enable_sky)
        if b.samples_per_frame:
            p.samples_per_frame = int(b.samples_per_frame)
        if b.max_bounces is not None:
            p.max_bounces = int(b.max_bounces)
        p.schedule, p.accum_mode, p.collect_stats = int(b.schedule), int(b.accum_mode), int(bool(b.collect_stats))
        if b.tile_rows is not None:  # (first_row, rows_per_band[, band_stride])
            p.tile_row_begin, p.tile_row_count = int(b.tile_rows[0]), int(b.tile_rows[1])
            p.tile_row_stride = int(b.tile_rows[2]) if len(b.tile_rows) > 2 else 0
        s = b.scene_description.handle
        h = lambda t: t.handle if t is not None else None
        if self.pipeline.kind == PATHTRACE:
            rc = self._lib.solb_trace_pathtrace(s, ctypes.byref(b.uniforms), ctypes.byref(p), h(b.accum_target), h(b.render_target))
        elif self.pipeline.kind == AO:
            rc = self._lib.solb_trace_ao(s, ctypes.byref(b.uniforms), ctypes.byref(p), h(b.accum_target))
        else:
            rc = self._lib.solb_trace_debug(s, ctypes.byref(b.uniforms), h(b.render_target), h(b.ids_target), None)
        N.check(rc, self.context.handle)


def resolve_sum(context, sum_target, accum_out=None, render=None):
    """Multi-GPU resolve after the reduce: accum = sum.xyz / sum.w, render = gamma 2.2 rgba8."""
    h = lambda t: t.handle if t is not None else None
    N.check(N.lib().solb_resolve_sum(context.handle, sum_target.handle, h(accum_out), h(render)), context.handle)
