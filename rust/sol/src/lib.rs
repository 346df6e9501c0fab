//! `sol` — the ray-tracing path of num3ric/sol-rs on top of libsolb.so (CUDA, sm_100a).
//!
//! What the three ray-tracing examples of the reference touch, under the reference's own paths:
//!   sol::scene::{load_scene, Scene, Mesh, ModelVertex, MaterialInfo, PrimitiveSection, BufferPart, Camera}
//!   sol::ray::{SceneDescription, SceneInstance, PipelineInfo, Pipeline, ShaderBindingTableInfo, ShaderBindingTable}
//!   sol::{Context, Image2d, ImageInfo, SceneUniforms}     (crate root, like `pub use crate::context::*` etc. in src/lib.rs:26-35)
//!   sol::comm::Communicator                                (new: the multi-GPU exchange of SURVEY 8e)
//! Windowing, swapchain, render passes, descriptors and the rasterisation examples are out of scope.
//! Error behaviour is the reference's: no `Result` on this path, a failed call panics (src/scene/mod.rs:140,
//! src/ray/pipeline.rs:105-115).
pub mod comm;
mod context;
pub mod ffi;
mod image2d;
pub mod ray;
pub mod scene;

pub use crate::context::*;
pub use crate::image2d::*;
pub use glam;

/// examples/5-pathtrace.rs:7-17 (same fields, same order, 400 bytes with the tail padding).
pub type SceneUniforms = ffi::SolbSceneUniforms;

impl SceneUniforms {
    /// examples/5-pathtrace.rs:19-32: `SceneUniforms::from(camera, uvec3(width, height, elapsed_ticks))`
    pub fn from(camera: &scene::Camera, frame: glam::UVec3) -> SceneUniforms {
        let view = camera.view_matrix();
        let projection = camera.perspective_matrix();
        SceneUniforms {
            model: glam::Mat4::IDENTITY.to_cols_array(),
            view: view.to_cols_array(),
            view_inverse: view.inverse().to_cols_array(),
            projection: projection.to_cols_array(),
            projection_inverse: projection.inverse().to_cols_array(),
            model_view_projection: (projection * view).to_cols_array(),
            frame: [frame.x, frame.y, frame.z],
            _pad: 0,
        }
    }
}
