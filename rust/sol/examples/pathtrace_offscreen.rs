//! examples/5-pathtrace.rs of the reference without the window: same setup() / render() call sequence
//! (load_scene, Camera, SceneDescription::from_scene, Pipeline, ShaderBindingTable, per frame SceneUniforms::from +
//! tlas_regenerate + cmd_trace_rays), the present blit replaced by a readback + PNG.
//!   cargo run --release --example pathtrace_offscreen -- assets/models/tunnel.gltf --sky 64 out.png
use sol::glam::{uvec3, vec2};
use sol::ray::{self, TraceBindings};
use sol::{scene, Context, Fence, HostBuffer, Image2d, ImageFormat, ImageInfo, SceneUniforms};
use std::path::PathBuf;

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let model = PathBuf::from(args.get(1).map(|s| s.as_str()).unwrap_or("assets/models/cornell.gltf"));
    let sky = args.iter().any(|a| a == "--sky"); // examples/5-pathtrace.rs:371-380
    let frames: u32 = args.iter().filter_map(|a| a.parse().ok()).next().unwrap_or(8);
    let out = args.iter().find(|a| a.ends_with(".png")).cloned().unwrap_or_else(|| "pathtrace.png".to_string());
    let (width, height) = (1920u32, 1080u32);

    // setup(): examples/5-pathtrace.rs:173-292
    let context = Context::new(0);
    let scene = scene::load_scene(context.clone(), &model);
    let mut camera = scene.camera.unwrap_or_else(|| scene::Camera::new(vec2(width as f32, height as f32)));
    camera.set_window_size(vec2(width as f32, height as f32));
    let mut scene_description = ray::SceneDescription::from_scene(context.clone(), &scene);
    let accum_target = Image2d::new(context.clone(), ImageInfo { width, height, format: ImageFormat::Rgba32f });
    let render_target = Image2d::new(context.clone(), ImageInfo { width, height, format: ImageFormat::Rgba8 });
    let pipeline = ray::Pipeline::new(
        context.clone(),
        ray::PipelineInfo::default()
            .shader(PathBuf::from("assets/glsl/pathtrace.rgen"), ray::SHADER_STAGE_RAYGEN_KHR)
            .shader(PathBuf::from("assets/glsl/pathtrace.rmiss"), ray::SHADER_STAGE_MISS_KHR)
            .shader(PathBuf::from("assets/glsl/pathtrace.rchit"), ray::SHADER_STAGE_CLOSEST_HIT_KHR)
            .specialization(&(sky as u32), 0)
            .name("pathtrace".to_string()),
    );
    let sbt = ray::ShaderBindingTable::new(context.clone(), pipeline.handle(), ray::ShaderBindingTableInfo::default().raygen(0).miss(1).hitgroup(2));

    // render(): examples/5-pathtrace.rs:294-369, once per frame.  Two frames in flight like the reference's renderer (one fence
    // per swapchain image, src/renderer.rs:72-81, 188): frame f's blit (here: the copy into a page-locked host buffer)
    // overlaps frame f + 1's tracing, and a slot's buffer is touched again only after its fence.
    let fences = [Fence::new(context.clone()), Fence::new(context.clone())];
    let mut present = [HostBuffer::new(context.clone(), render_target.size_bytes()), HostBuffer::new(context.clone(), render_target.size_bytes())];
    for frame in 0..frames {
        let slot = (frame & 1) as usize;
        fences[slot].wait(); // wait_for_and_reset_fence
        let uniforms = SceneUniforms::from(&camera, uvec3(width, height, frame));
        scene_description.tlas_regenerate(());
        let mut bindings = TraceBindings::new(&scene_description, &uniforms);
        bindings.accum_target = Some(&accum_target);
        bindings.render_target = Some(&render_target);
        sbt.cmd_trace_rays(&bindings, (width, height, 1));
        render_target.readback_async(&mut present[slot]); // replaces cmd_blit_to(present image)
        fences[slot].signal(); // queue_submit(.., in_flight_fence)
    }
    let last = ((frames.max(1) - 1) & 1) as usize;
    fences[0].wait();
    fences[1].wait();
    let pixels = present[last].as_slice().to_vec();
    let stats = context.stats();
    println!("{} frames, {} rays, {} kernel launches", frames, stats.rays, stats.kernel_launches);
    image::save_buffer(&out, &pixels, width, height, image::ColorType::Rgba8).unwrap();
}
