// pathtrace_offscreen.cpp — examples/5-pathtrace.rs of the reference without the window: the same setup() /
// render() call sequence through the C++ host mirror (sol.hpp), frames read back instead of presented.
//   pathtrace_offscreen --model models/cornell.gltf [--sky] [--frames 8] [--size 512x512] [--bounces 32] [--textures]
//                       [--debug] [--two-level] [--out frame.ppm]
// Prints one line: frames, ms/frame, Mrays/s, and an FNV-1a checksum of the final rgba8 frame.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../sol_rs_b200/csrc/host/sol.hpp"

using namespace sol;

int main(int argc, char **argv) {
    std::string model, out;
    bool enable_sky = false, debug = false, two_level = false, textures = false;
    uint32_t frames = 8, w = 1280, h = 720, bounces = 0;  // 1280x720: examples/5-pathtrace.rs:374
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--model" && i + 1 < argc) model = argv[++i];       // examples/5-pathtrace.rs:120-125
        else if (a == "--sky") enable_sky = true;                     // :220
        else if (a == "--textures") textures = true;                  // beyond the reference: bind the base-colour textures
        else if (a == "--debug") debug = true;
        else if (a == "--two-level") two_level = true;  // TLAS over object-space BLASes instead of the flattened hierarchy
        else if (a == "--frames" && i + 1 < argc) frames = (uint32_t)atoi(argv[++i]);
        else if (a == "--bounces" && i + 1 < argc) bounces = (uint32_t)atoi(argv[++i]);
        else if (a == "--out" && i + 1 < argc) out = argv[++i];
        else if (a == "--size" && i + 1 < argc) { if (sscanf(argv[++i], "%ux%u", &w, &h) != 2) { fprintf(stderr, "bad --size\n"); return 2; } }
    }
    if (model.empty()) { fprintf(stderr, "no gltf file given\n"); return 2; }  // the reference panics here
    try {
        auto context = Context::create(0);
        // ---- setup(): examples/5-pathtrace.rs:118-248 ----
        auto path = util::find_asset(model);
        if (!path) path = model;
        scene::Scene scene = scene::load_scene(context, *path);
        ray::SceneDescription scene_description = ray::SceneDescription::from_scene(context, scene);
        if (textures) scene_description.set_textures(scene);
        if (two_level) {
            scene_description.set_accel_mode(SOLB_ACCEL_TWO_LEVEL);
            scene_description.accel_build();
        }
        scene::Camera camera = scene.camera ? *scene.camera : scene::Camera(Vec2{ (float)w, (float)h });
        if (debug) { camera = scene::Camera(Vec2{ (float)w, (float)h }); camera.look_at({ 5, 5, 5 }, { 0, 0, 0 }, { 0, -1, 0 }); }
        else camera.set_window_size(Vec2{ (float)w, (float)h });
        const uint32_t spec[1] = { enable_sky ? 1u : 0u };
        const std::string stem = debug ? "debug" : "pathtrace";
        ray::Pipeline pipeline(context, ray::PipelineInfo()
                                            .shader("glsl/" + stem + ".rgen", ray::ShaderStage::RAYGEN_KHR)
                                            .shader("glsl/" + stem + ".rmiss", ray::ShaderStage::MISS_KHR)
                                            .shader("glsl/" + stem + ".rchit", ray::ShaderStage::CLOSEST_HIT_KHR)
                                            .specialization(spec, 1, 0)
                                            .name("AO_mat"));
        ray::ShaderBindingTable sbt(context, pipeline, ray::ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2));
        Image2d accum_target(context, w, h, SOLB_FORMAT_RGBA32F);
        Image2d render_target(context, w, h, SOLB_FORMAT_RGBA8);
        const uint32_t accumulation_start_frame = 0;
        std::vector<uint8_t> frame(render_target.size_bytes());
        // ---- render(): examples/5-pathtrace.rs:294-369, elapsed_ticks = 0, 1, 2, ... (src/lib.rs:234) ----
        const SolbStats s0 = context->stats();
        const auto t0 = std::chrono::steady_clock::now();
        for (uint32_t elapsed_ticks = 0; elapsed_ticks < frames; elapsed_ticks++) {
            const SceneUniforms uniforms = SceneUniforms::from(camera, { w, h, elapsed_ticks });
            scene_description.tlas_regenerate();
            ray::TraceBindings b;
            b.scene_description = &scene_description;
            b.uniforms = &uniforms;
            b.accumulation_start_frame = accumulation_start_frame;
            b.accum_target = debug ? nullptr : &accum_target;
            b.render_target = &render_target;
            b.overrides.max_bounces = bounces;
            sbt.cmd_trace_rays(b, { w, h, 1 });
            render_target.readback(frame.data(), frame.size());  // blit + present
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const SolbStats s1 = context->stats();
        uint64_t hash = 1469598103934665603ull;
        for (uint8_t b : frame) { hash ^= b; hash *= 1099511628211ull; }
        printf("{\"model\": \"%s\", \"size\": [%u, %u], \"frames\": %u, \"ms_per_frame\": %.3f, \"Mrays_s\": %.1f, \"rays\": %llu, \"fnv1a\": \"%016llx\"}\n",
               model.c_str(), w, h, frames, 1e3 * sec / frames, (double)(s1.rays - s0.rays) / sec / 1e6,
               (unsigned long long)(s1.rays - s0.rays), (unsigned long long)hash);
        if (!out.empty()) {
            FILE *f = fopen(out.c_str(), "wb");
            if (!f) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
            fprintf(f, "P6\n%u %u\n255\n", w, h);
            for (size_t p = 0; p < (size_t)w * h; p++) fwrite(&frame[4 * p], 1, 3, f);
            fclose(f);
        }
    } catch (const Error &e) {
        fprintf(stderr, "sol::Error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
