//! Drop-in for sol::ray (src/ray/{mod,pipeline,sbt}.rs) on top of libsolb.  Same type names, constructor
//! arguments and call order as the reference; Vulkan handles the reference passes through (command buffers,
//! pipeline layouts) are accepted and ignored.  A failed C call panics, like the reference's unwrap()s.
use crate::ffi::*;
use crate::scene::{Mesh, Scene};
use std::ffi::CStr;
use std::sync::Arc;

pub struct Context { pub(crate) raw: *mut solb_ctx }
impl Context {
    pub fn new(device: i32) -> Arc<Context> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { solb_ctx_create(device, std::ptr::null_mut(), &mut raw) };
        assert!(rc == 0, "solb_ctx_create: {}", last_error(std::ptr::null_mut()));
        Arc::new(Context { raw })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { solb_ctx_destroy(self.raw); } } }

pub(crate) fn last_error(ctx: *mut solb_ctx) -> String {
    unsafe { CStr::from_ptr(solb_last_error(ctx)).to_string_lossy().into_owned() }
}
fn check(ctx: &Context, rc: i32) { if rc != 0 { panic!("libsolb: {}", last_error(ctx.raw)); } }

pub use crate::ffi::SolbSceneInstance as SceneInstance;

pub struct SceneDescription { context: Arc<Context>, raw: *mut solb_scene }

impl SceneDescription {
    /// src/ray/mod.rs:50-57
    pub fn from_scene(context: Arc<Context>, scene: &Scene) -> Self {
        let meshes = scene.meshes.iter().collect::<Vec<_>>();
        let transforms = scene.meshes.iter().map(|m| m.transform).collect::<Vec<_>>();
        Self::from_meshes(context, meshes, transforms, Some(&scene.materials))
    }
    /// src/ray/mod.rs:59-156: one BLAS + one instance per primitive section, then the TLAS.
    pub fn from_meshes(context: Arc<Context>, meshes: Vec<&Mesh>, mesh_transforms: Vec<glam::Mat4>,
                       materials: Option<&Vec<crate::scene::MaterialInfo>>) -> Self {
        let sections: Vec<Vec<SolbSection>> = meshes.iter().map(|m| m.primitive_sections.iter().map(|p| SolbSection {
            first_vertex: p.get_vertex_offset(), n_vertices: p.get_vertex_count(),
            first_index: p.get_indices().map(|i| i.offset as u32).unwrap_or(0),
            n_indices: p.get_indices().map(|i| i.element_count as u32).unwrap_or(0),
            material_index: p.material_index.unwrap() as u32,           // reference: src/scene/mod.rs:65
        }).collect()).collect();
        let descs: Vec<SolbMeshDesc> = meshes.iter().zip(&sections).zip(&mesh_transforms).map(|((m, s), t)| SolbMeshDesc {
            vertices: m.vertices.as_ptr() as *const SolbModelVertex, n_vertices: m.vertices.len() as u32,
            indices: m.indices.as_ptr(), n_indices: m.indices.len() as u32,
            sections: s.as_ptr(), n_sections: s.len() as u32, transform: t.to_cols_array(),
        }).collect();
        let (mp, mn) = materials.map(|m| (m.as_ptr() as *const SolbMaterialInfo, m.len() as u32)).unwrap_or((std::ptr::null(), 0));
        let mut raw = std::ptr::null_mut();
        check(&context, unsafe { solb_scene_create(context.raw, descs.as_ptr(), descs.len() as u32, mp, mn, &mut raw) });
        check(&context, unsafe { solb_accel_build(raw) });               // BLAS::new x n + TLAS::new
        SceneDescription { context, raw }
    }
    /// src/ray/mod.rs:162-167
    pub fn blas_transform(&mut self, transform: glam::Mat4, index: usize) {
        check(&self.context, unsafe { solb_instance_set_transform(self.raw, index as u32, transform.to_cols_array().as_ptr()) });
    }
    pub fn blas_transforms(&mut self, transforms: &[glam::Mat4]) {
        transforms.iter().enumerate().for_each(|(i, t)| self.blas_transform(*t, i));
    }
    /// src/ray/mod.rs:178-181 (the command buffer argument is accepted for source compatibility)
    pub fn tlas_regenerate<C>(&mut self, _cmd: C) { check(&self.context, unsafe { solb_tlas_regenerate(self.raw) }); }
    /// Beyond the reference (its TODO at src/ray/mod.rs:122): one more instance of the BLAS `source_instance` uses.
    /// Returns the new gl_InstanceID; takes effect at the next `accel_build` / `tlas_regenerate`.
    pub fn add_instance(&mut self, source_instance: usize, transform: glam::Mat4, material_index: u32) -> u32 {
        let mut id = 0u32;
        check(&self.context, unsafe { solb_scene_add_instance(self.raw, source_instance as u32, transform.to_cols_array().as_ptr(), material_index, &mut id) });
        id
    }
    /// 0 = flattened (default), 1 = two-level: TLAS over shared BLASes, `tlas_regenerate` rebuilds the TLAS only.
    pub fn set_accel_mode(&mut self, mode: u32) { check(&self.context, unsafe { solb_scene_set_accel_mode(self.raw, mode) }); }
    pub fn accel_build(&mut self) { check(&self.context, unsafe { solb_accel_build(self.raw) }); }
    /// src/ray/mod.rs:190-192
    pub fn update(&mut self) { check(&self.context, unsafe { solb_scene_update(self.raw) }); }
    pub(crate) fn raw(&self) -> *mut solb_scene { self.raw }
}
impl Drop for SceneDescription { fn drop(&mut self) { unsafe { solb_scene_destroy(self.raw); } } }

#[derive(Clone, Copy, PartialEq)] pub enum Kind { Pathtrace, Ao, Debug }

/// src/ray/pipeline.rs:5-52
#[derive(Default)]
pub struct PipelineInfo { shaders: Vec<(std::path::PathBuf, u32)>, spec: Vec<u32>, spec_id: u32, name: String }
impl PipelineInfo {
    pub fn layout<L>(self, _layout: L) -> Self { self }
    pub fn shader(mut self, path: std::path::PathBuf, stage: u32) -> Self { self.shaders.push((path, stage)); self }
    pub fn specialization(mut self, data: &[u32], constant_id: u32) -> Self { self.spec = data.to_vec(); self.spec_id = constant_id; self }
    pub fn name(mut self, name: String) -> Self { self.name = name; self }
}
pub const STAGE_RAYGEN: u32 = 0x100; // vk::ShaderStageFlags::RAYGEN_KHR

/// src/ray/pipeline.rs:61-122: kernels are compiled ahead of time; "new" selects the family by raygen file name.
pub struct Pipeline { pub(crate) kind: Kind, pub(crate) enable_sky: bool }
impl Pipeline {
    pub fn new(_context: Arc<Context>, info: PipelineInfo) -> Self {
        let rgen = info.shaders.iter().find(|(_, s)| *s == STAGE_RAYGEN).expect("no raygen stage");
        let kind = match rgen.0.file_name().and_then(|f| f.to_str()) {
            Some("pathtrace.rgen") => Kind::Pathtrace, Some("ao.rgen") => Kind::Ao, Some("debug.rgen") => Kind::Debug,
            other => panic!("no CUDA kernel family for raygen shader {:?}", other),
        };
        Pipeline { kind, enable_sky: info.spec_id == 0 && info.spec.first().copied().unwrap_or(0) != 0 }
    }
}

/// What the reference binds through descriptor sets + push constants (examples/5-pathtrace.rs:297-356).
pub struct TraceBindings<'a> {
    pub scene_description: &'a SceneDescription, pub uniforms: &'a SolbSceneUniforms, pub accumulation_start_frame: u32,
    pub accum_target: Option<&'a crate::Image2d>, pub render_target: Option<&'a crate::Image2d>, pub ids_target: Option<&'a crate::Image2d>,
}

/// src/ray/sbt.rs:58-181
pub struct ShaderBindingTable { context: Arc<Context>, kind: Kind, enable_sky: bool }
impl ShaderBindingTable {
    pub fn new(context: Arc<Context>, pipeline: &Pipeline) -> Self { ShaderBindingTable { context, kind: pipeline.kind, enable_sky: pipeline.enable_sky } }
    /// src/ray/sbt.rs:167-180
    pub fn cmd_trace_rays(&self, b: &TraceBindings, _extent: (u32, u32, u32)) {
        let mut p = SolbTraceParams::default();
        unsafe { solb_trace_params_default(&mut p, if self.kind == Kind::Ao { 1 } else { 0 }) };
        p.accum_start_frame = b.accumulation_start_frame as i32;
        p.enable_sky = self.enable_sky as u32;
        let t = |x: Option<&crate::Image2d>| x.map(|i| i.raw).unwrap_or(std::ptr::null_mut());
        let rc = unsafe { match self.kind {
            Kind::Pathtrace => solb_trace_pathtrace(b.scene_description.raw(), b.uniforms, &p, t(b.accum_target), t(b.render_target)),
            Kind::Ao => solb_trace_ao(b.scene_description.raw(), b.uniforms, &p, t(b.accum_target)),
            Kind::Debug => solb_trace_debug(b.scene_description.raw(), b.uniforms, t(b.render_target), t(b.ids_target), std::ptr::null_mut()),
        } };
        check(&self.context, rc);
    }
}
