//! sol::ray on top of libsolb (src/ray/{mod,pipeline,sbt,acceleration}.rs): same type names, constructor arguments and call
//! order as the reference.  Vulkan objects the reference threads through (command buffers, pipeline layouts, vk::Pipeline)
//! are generic parameters that are accepted and ignored, so `sbt.cmd_trace_rays(...)`-style call sites keep their shape.
//! A failed C call panics, like the reference's unwrap()s.
use crate::ffi::*;
use crate::scene::{MaterialInfo, Mesh, Scene};
use crate::{Context, Image2d};
use glam::Mat4;
use std::path::PathBuf;
use std::sync::Arc;

/// src/ray/mod.rs:16-24 (what `instances[gl_InstanceID]` is to the shaders)
pub type SceneInstance = SolbSceneInstance;

/// How the acceleration structure is organised (beyond the reference).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum AccelMode {
    /// instance transforms baked into world-space triangles, ONE hierarchy (default)
    Flat,
    /// TLAS of instances over shared object-space BLASes: `tlas_regenerate` rebuilds the TLAS only
    TwoLevel,
}

/// src/ray/mod.rs:37-48.  BLAS / TLAS / instance + descriptor buffers all live behind the one libsolb scene handle.
pub struct SceneDescription {
    context: Arc<Context>,
    raw: *mut solb_scene,
}

impl SceneDescription {
    /// src/ray/mod.rs:50-57
    pub fn from_scene(context: Arc<Context>, scene: &Scene) -> Self {
        let meshes = scene.meshes.iter().collect::<Vec<_>>();
        let transforms = scene.meshes.iter().map(|m| m.transform).collect::<Vec<_>>();
        Self::from_meshes(context, meshes, transforms, Some(&scene.materials))
    }

    /// src/ray/mod.rs:59-156: one BLAS and one instance per primitive section (instance id = running count), then the TLAS.
    /// The reference takes `Option<&Buffer>` for the materials; the host array is what libsolb copies.
    pub fn from_meshes(context: Arc<Context>, meshes: Vec<&Mesh>, mesh_transforms: Vec<Mat4>, materials: Option<&Vec<MaterialInfo>>) -> Self {
        let mut sd = Self::create(context, &meshes, &mesh_transforms, materials);
        sd.accel_build(); // BLAS::new x n + TLAS::new
        sd
    }

    /// from_scene + every further glTF node of a mesh as an instance of the same BLASes, in `mode`.
    pub fn from_scene_instanced(context: Arc<Context>, scene: &Scene, mode: AccelMode) -> Self {
        let meshes = scene.meshes.iter().collect::<Vec<_>>();
        let transforms = scene.meshes.iter().map(|m| m.transform).collect::<Vec<_>>();
        let mut sd = Self::create(context, &meshes, &transforms, Some(&scene.materials));
        sd.set_accel_mode(mode);
        let mut first_instance = 0usize; // instance id of the mesh's first section
        for mesh in &scene.meshes {
            for t in &mesh.extra_instance_transforms {
                for (k, section) in mesh.primitive_sections.iter().enumerate() {
                    sd.add_instance(first_instance + k, *t, section.get_material_index() as u32);
                }
            }
            first_instance += mesh.primitive_sections.len();
        }
        sd.accel_build();
        sd
    }

    fn create(context: Arc<Context>, meshes: &[&Mesh], mesh_transforms: &[Mat4], materials: Option<&Vec<MaterialInfo>>) -> Self {
        assert_eq!(meshes.len(), mesh_transforms.len());
        let sections: Vec<Vec<SolbSection>> = meshes
            .iter()
            .map(|m| {
                m.primitive_sections
                    .iter()
                    .map(|p| SolbSection {
                        first_vertex: p.get_vertex_offset(),
                        n_vertices: p.get_vertex_count(),
                        first_index: p.get_indices().map(|i| i.offset as u32).unwrap_or(0),
                        n_indices: p.get_indices().map(|i| i.element_count as u32).unwrap_or(0),
                        material_index: p.get_material_index() as u32, // unwrap(): src/scene/mod.rs:65
                    })
                    .collect()
            })
            .collect();
        let descs: Vec<SolbMeshDesc> = meshes
            .iter()
            .zip(&sections)
            .zip(mesh_transforms)
            .map(|((m, s), t)| SolbMeshDesc {
                vertices: m.vertices.as_ptr(),
                n_vertices: m.vertices.len() as u32,
                indices: m.indices.as_ptr(),
                n_indices: m.indices.len() as u32,
                sections: s.as_ptr(),
                n_sections: s.len() as u32,
                transform: t.to_cols_array(),
            })
            .collect();
        let (mat_ptr, mat_len) = match materials {
            Some(m) => (m.as_ptr(), m.len() as u32),
            None => (std::ptr::null(), 0),
        };
        let mut raw = std::ptr::null_mut();
        context.check(unsafe { solb_scene_create(context.raw, descs.as_ptr(), descs.len() as u32, mat_ptr, mat_len, &mut raw) });
        SceneDescription { context, raw }
    }

    /// src/ray/mod.rs:162-167
    pub fn blas_transform(&mut self, transform: Mat4, index: usize) {
        self.context.check(unsafe { solb_instance_set_transform(self.raw, index as u32, transform.to_cols_array().as_ptr()) });
    }

    /// src/ray/mod.rs:169-176
    pub fn blas_transforms(&mut self, transforms: &[Mat4]) {
        for (index, transform) in transforms.iter().enumerate() {
            self.blas_transform(*transform, index);
        }
    }

    /// src/ray/mod.rs:178-181 (the command buffer is accepted for source compatibility).  No-op while no transform changed.
    pub fn tlas_regenerate<C>(&mut self, _cmd: C) {
        self.context.check(unsafe { solb_tlas_regenerate(self.raw) });
    }

    /// src/ray/mod.rs:190-192
    pub fn update(&mut self) {
        self.context.check(unsafe { solb_scene_update(self.raw) });
    }

    /// The instance array as the shaders see it (replaces get_instances_buffer, src/ray/mod.rs:186-188).
    pub fn get_instances(&self) -> Vec<SceneInstance> {
        let mut n = 0u32;
        self.context.check(unsafe { solb_scene_instance_count(self.raw, &mut n) });
        let zero = SceneInstance { id: 0, texture_offset: 0, padding: [0.0; 2], transform: [0.0; 16], transform_it: [0.0; 16] };
        let mut out = vec![zero; n as usize];
        self.context.check(unsafe { solb_scene_get_instances(self.raw, out.as_mut_ptr(), n) });
        out
    }

    /// primitive_count of every instance's BLAS (src/ray/acceleration.rs:183), instance order
    pub fn get_instance_triangle_counts(&self) -> Vec<u32> {
        let mut n = 0u32;
        self.context.check(unsafe { solb_scene_instance_count(self.raw, &mut n) });
        let mut out = vec![0u32; n as usize];
        self.context.check(unsafe { solb_scene_instance_triangles(self.raw, out.as_mut_ptr(), n) });
        out
    }

    /// What replaces `tlas()` / `blas()` (src/ray/mod.rs:158-160,183-185): sizes and quality of the built structure.
    pub fn accel_info(&self) -> SolbAccelInfo {
        let mut info = SolbAccelInfo::default();
        self.context.check(unsafe { solb_accel_info(self.raw, &mut info) });
        info
    }

    /// Beyond the reference (its TODO at src/ray/mod.rs:122): one more instance of the BLAS `source_instance` uses.
    /// Returns the new gl_InstanceID; takes effect at the next `accel_build` / `tlas_regenerate`.
    pub fn add_instance(&mut self, source_instance: usize, transform: Mat4, material_index: u32) -> u32 {
        let mut id = 0u32;
        self.context.check(unsafe {
            solb_scene_add_instance(self.raw, source_instance as u32, transform.to_cols_array().as_ptr(), material_index, &mut id)
        });
        id
    }

    /// Beyond the reference: bind base-colour textures (rgba8 rows top first, sRGB colour) and, per material, its
    /// texture (`None`: untextured).  Every instance's `texture_offset` (src/ray/mod.rs:20) becomes the texture of its material.
    pub fn set_textures(&mut self, textures: &[crate::scene::Texture], material_textures: &[Option<u32>]) {
        let descs: Vec<SolbTextureDesc> = textures
            .iter()
            .map(|t| SolbTextureDesc { rgba8: t.rgba8.as_ptr(), width: t.width, height: t.height, wrap_s: t.wrap_s, wrap_t: t.wrap_t, srgb: 1, _pad: 0 })
            .collect();
        let mt: Vec<u32> = material_textures.iter().map(|t| t.unwrap_or(SOLB_NO_TEXTURE)).collect();
        self.context.check(unsafe {
            solb_scene_set_textures(self.raw, descs.as_ptr(), descs.len() as u32, mt.as_ptr(), mt.len() as u32)
        });
    }

    pub fn set_accel_mode(&mut self, mode: AccelMode) {
        let m = if mode == AccelMode::TwoLevel { SOLB_ACCEL_TWO_LEVEL } else { SOLB_ACCEL_FLAT };
        self.context.check(unsafe { solb_scene_set_accel_mode(self.raw, m) });
    }

    pub fn accel_build(&mut self) {
        self.context.check(unsafe { solb_accel_build(self.raw) });
    }

    /// traceRayEXT for host rays {ox,oy,oz,tmin, dx,dy,dz,tmax} -> ({instance, primitive, bits(u), bits(v)}, t)
    pub fn trace_rays(&self, rays: &[[f32; 8]]) -> (Vec<[u32; 4]>, Vec<f32>) {
        let mut hits = vec![[0u32; 4]; rays.len()];
        let mut t = vec![0f32; rays.len()];
        self.context.check(unsafe {
            solb_trace_rays(self.raw, rays.as_ptr() as *const f32, rays.len() as u32, hits.as_mut_ptr() as *mut u32, t.as_mut_ptr())
        });
        (hits, t)
    }

    pub(crate) fn raw(&self) -> *mut solb_scene {
        self.raw
    }
}

impl Drop for SceneDescription {
    fn drop(&mut self) {
        unsafe {
            solb_scene_destroy(self.raw);
        }
    }
}

/// Kernel family a pipeline selects (the raygen stage's file name).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum Kind {
    Pathtrace,
    Ao,
    Debug,
}

/// vk::ShaderStageFlags bits the examples pass to PipelineInfo::shader
pub const SHADER_STAGE_RAYGEN_KHR: u32 = 0x100;
pub const SHADER_STAGE_CLOSEST_HIT_KHR: u32 = 0x400;
pub const SHADER_STAGE_MISS_KHR: u32 = 0x800;

/// src/ray/pipeline.rs:5-52
#[derive(Default)]
pub struct PipelineInfo {
    shaders: Vec<(PathBuf, u32)>,
    specialization_data: Vec<u32>,
    specialization_id: u32,
    name: String,
}

impl PipelineInfo {
    /// the reference passes a vk::PipelineLayout: no counterpart here
    pub fn layout<L>(self, _layout: L) -> Self {
        self
    }
    pub fn shader(mut self, path: PathBuf, stage: u32) -> Self {
        self.shaders.push((path, stage));
        self
    }
    /// src/ray/pipeline.rs:38-47: constant 0 of the pathtrace pipeline = ENABLE_SKYLIGHT (examples/5-pathtrace.rs:103)
    pub fn specialization<T: Copy + Into<u32>>(mut self, data: &T, constant_id: u32) -> Self {
        self.specialization_data = vec![(*data).into()];
        self.specialization_id = constant_id;
        self
    }
    pub fn name(mut self, name: String) -> Self {
        self.name = name;
        self
    }
}

/// src/ray/pipeline.rs:54-122: the kernels are compiled ahead of time; `new` selects the family by the raygen file name,
/// captures specialization constant 0 and loads the kernels (the counterpart of compiling the GLSL stages).
pub struct Pipeline {
    kind: Kind,
    enable_sky: bool,
}

impl Pipeline {
    pub fn new(context: Arc<Context>, info: PipelineInfo) -> Self {
        let rgen = info.shaders.iter().find(|(_, stage)| *stage == SHADER_STAGE_RAYGEN_KHR).expect("ray::Pipeline: no raygen stage");
        let kind = match rgen.0.file_name().and_then(|f| f.to_str()) {
            Some("pathtrace.rgen") => Kind::Pathtrace,
            Some("ao.rgen") => Kind::Ao,
            Some("debug.rgen") => Kind::Debug,
            other => panic!("ray::Pipeline: no CUDA kernel family for raygen shader {:?}", other),
        };
        context.preload();
        Pipeline { kind, enable_sky: info.specialization_id == 0 && info.specialization_data.first().copied().unwrap_or(0) != 0 }
    }
    /// stands in for `pipeline.handle()` (a vk::Pipeline in the reference) at the ShaderBindingTable::new call site
    pub fn handle(&self) -> &Pipeline {
        self
    }
    pub fn kind(&self) -> Kind {
        self.kind
    }
}

/// src/ray/sbt.rs:13-55
#[derive(Default)]
pub struct ShaderBindingTableInfo {
    pub raygen_indices: Vec<u64>,
    pub miss_indices: Vec<u64>,
    pub hit_group_indices: Vec<u64>,
}

impl ShaderBindingTableInfo {
    pub fn raygen(mut self, index: u64) -> Self {
        self.raygen_indices.push(index);
        self
    }
    pub fn miss(mut self, index: u64) -> Self {
        self.miss_indices.push(index);
        self
    }
    pub fn hitgroup(mut self, index: u64) -> Self {
        self.hit_group_indices.push(index);
        self
    }
}

/// Everything the reference binds through descriptor sets + the push constant before cmd_trace_rays
/// (examples/5-pathtrace.rs:297-356), plus the shader literals BASELINE.json's configs vary (SolbTraceParams).
pub struct TraceBindings<'a> {
    pub scene_description: &'a SceneDescription,
    pub uniforms: &'a crate::SceneUniforms,
    /// push.accum_start_frame (examples/5-pathtrace.rs:306-314)
    pub accumulation_start_frame: u32,
    /// binding "accumulation image" (pathtrace) / the AO image
    pub accum_target: Option<&'a Image2d>,
    pub render_target: Option<&'a Image2d>,
    pub ids_target: Option<&'a Image2d>,
    /// None = the reference's literals (8 spp, 32 bounces; ao: 4, 4)
    pub samples_per_frame: Option<u32>,
    pub max_bounces: Option<u32>,
    /// SOLB_ACCUM_MIX (reference) or SOLB_ACCUM_SUM (multi-GPU frames split)
    pub accum_mode: u32,
    /// (first row, rows per band, stride between bands) of a tile-split frame; None = whole image
    pub tile_rows: Option<(u32, u32, u32)>,
}

impl<'a> TraceBindings<'a> {
    pub fn new(scene_description: &'a SceneDescription, uniforms: &'a crate::SceneUniforms) -> Self {
        TraceBindings {
            scene_description,
            uniforms,
            accumulation_start_frame: 0,
            accum_target: None,
            render_target: None,
            ids_target: None,
            samples_per_frame: None,
            max_bounces: None,
            accum_mode: SOLB_ACCUM_MIX,
            tile_rows: None,
        }
    }
}

/// src/ray/sbt.rs:58-181
pub struct ShaderBindingTable {
    context: Arc<Context>,
    kind: Kind,
    enable_sky: bool,
}

impl ShaderBindingTable {
    /// src/ray/sbt.rs:71: `ShaderBindingTable::new(context, pipeline.handle(), info.raygen(0).miss(1).hitgroup(2))`
    pub fn new(context: Arc<Context>, pipeline: &Pipeline, info: ShaderBindingTableInfo) -> Self {
        if info.raygen_indices.len() != 1 || info.miss_indices.len() != 1 || info.hit_group_indices.len() != 1 {
            panic!("ShaderBindingTable: exactly one raygen / miss / hit group is supported");
        }
        ShaderBindingTable { context, kind: pipeline.kind, enable_sky: pipeline.enable_sky }
    }

    /// src/ray/sbt.rs:167-180: one launch of extent (width, height, 1) = the size of the bound targets.
    pub fn cmd_trace_rays(&self, b: &TraceBindings, _extent: (u32, u32, u32)) {
        let mut p = SolbTraceParams::default();
        unsafe { solb_trace_params_default(&mut p, if self.kind == Kind::Ao { 1 } else { 0 }) };
        p.accum_start_frame = b.accumulation_start_frame as i32;
        p.enable_sky = self.enable_sky as u32;
        p.accum_mode = b.accum_mode;
        if let Some(spp) = b.samples_per_frame {
            p.samples_per_frame = spp;
        }
        if let Some(mb) = b.max_bounces {
            p.max_bounces = mb;
        }
        if let Some((begin, count, stride)) = b.tile_rows {
            p.tile_row_begin = begin;
            p.tile_row_count = count;
            p.tile_row_stride = stride;
        }
        let raw = |t: Option<&Image2d>| t.map(|i| i.raw).unwrap_or(std::ptr::null_mut());
        let scene = b.scene_description.raw();
        let rc = unsafe {
            match self.kind {
                Kind::Pathtrace => solb_trace_pathtrace(scene, b.uniforms, &p, raw(b.accum_target), raw(b.render_target)),
                Kind::Ao => solb_trace_ao(scene, b.uniforms, &p, raw(b.accum_target)),
                Kind::Debug => solb_trace_debug(scene, b.uniforms, raw(b.render_target), raw(b.ids_target), std::ptr::null_mut()),
            }
        };
        self.context.check(rc);
    }
}

/// Multi-GPU resolve without a communicator: out = sum.xyz / sum.w with the reference's display transform.
pub fn resolve_sum(context: &Arc<Context>, sum: &Image2d, accum_out: Option<&Image2d>, render: Option<&Image2d>) {
    let raw = |t: Option<&Image2d>| t.map(|i| i.raw).unwrap_or(std::ptr::null_mut());
    context.check(unsafe { solb_resolve_sum(context.raw, sum.raw, raw(accum_out), raw(render)) });
}
