#!/usr/bin/env python
"""Numbers for the BASELINE.json configs that are not the headline line of bench.py:
  ao      configs[1]  4-ray-ao, Duck.gltf (ToyCar.glb is missing upstream), 1920x1080, frame 0..3
  4k      configs[3]  5-pathtrace tunnel --sky 3840x2160 cap 8 (single GPU here; multi-GPU through bench.py --gpus)
  synth   configs[4]  synthetic instanced scene: build ms + 1080p path trace (size via --blas / --grid)
  build   build-time sweep of the shipped scenes + synthetic sizes
Each prints one JSON line.  usage: python tools/bench_configs.py ao 4k synth build [--blas 1000 --grid 100]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="+")
    ap.add_argument("--blas", type=int, default=1000)
    ap.add_argument("--grid", type=int, default=100)
    ap.add_argument("--frames", type=int, default=4)
    args = ap.parse_args()
    import torch

    import sol_rs_b200 as sol
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene, synth

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = sol.Context(0, stream.cuda_stream)

    def pt_sbt(sky):
        pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/pathtrace.rgen", ray.RAYGEN_KHR).shader("glsl/pathtrace.rmiss", ray.MISS_KHR)
                            .shader("glsl/pathtrace.rchit", ray.CLOSEST_HIT_KHR).specialization([int(sky)], 0))
        return ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))

    def timed_frames(fn, n, warm=2):
        for f in range(warm):
            fn(f)
        ctx.reset_stats()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for f in range(n):
            fn(warm + f)
        e1.record(stream)
        torch.cuda.synchronize()
        st = ctx.stats()
        ms = e0.elapsed_time(e1)
        return {"ms_per_frame": ms / n, "Mrays_s": st.rays / (ms * 1e-3) / 1e6, "rays_per_frame": st.rays / n,
                "rays_per_path": st.rays / max(st.paths, 1), "launches_per_frame": st.kernel_launches / n}

    def pathtrace(sd, cam, w, h, sky, mb, frames, schedule=N.SCHEDULE_AUTO):
        cam.set_window_size((w, h))
        accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
        render = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
        sbt = pt_sbt(sky)
        return timed_frames(lambda f: sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, render,
                                                                            max_bounces=mb, schedule=schedule), (w, h, 1)), frames)

    for what in args.what:
        if what == "ao":
            from helpers import load_blue_noise

            w, h = 1920, 1080
            sc = scene.load_scene(ctx, os.path.join(ROOT, "assets/models/Duck.gltf"))
            sd = ray.SceneDescription.from_scene(ctx, sc)
            cam = scene.Camera((w, h))
            cam.look_at((4, 1, 4), (0, 0.5, 0), (0, -1, 0))  # examples/4-ray-ao.rs:89-90
            ctx.set_blue_noise(load_blue_noise())
            img = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
            pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/ao.rgen", ray.RAYGEN_KHR).shader("glsl/ao.rmiss", ray.MISS_KHR)
                                .shader("glsl/ao.rchit", ray.CLOSEST_HIT_KHR))
            sbt = ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))
            r = timed_frames(lambda f: sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), img, None), (w, h, 1)), args.frames)
            print(json.dumps({"config": "4-ray-ao Duck.gltf (ToyCar.glb missing) 1920x1080, 4 samples x (1 primary + <=4 AO)", **r}), flush=True)
        elif what == "4k":
            sc = scene.load_scene(ctx, os.path.join(ROOT, "assets/models/tunnel.gltf"))
            sd = ray.SceneDescription.from_scene(ctx, sc)
            r = pathtrace(sd, sc.camera, 3840, 2160, True, 8, args.frames)
            print(json.dumps({"config": "5-pathtrace tunnel.gltf --sky 3840x2160 cap 8, 1 GPU", **r}), flush=True)
        elif what == "converge":
            # configs[2] in full: 512 frames x 8 spp = 4096 spp at 1920x1080, readback of the final frame only
            sc = scene.load_scene(ctx, os.path.join(ROOT, "assets/models/tunnel.gltf"))
            sd = ray.SceneDescription.from_scene(ctx, sc)
            cam = sc.camera
            w, h = 1920, 1080
            cam.set_window_size((w, h))
            accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
            render = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
            sbt = pt_sbt(True)
            ctx.reset_stats()
            torch.cuda.synchronize()
            t0 = time.time()
            for f in range(512):
                sd.tlas_regenerate()
                sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, render, max_bounces=8), (w, h, 1))
            img = render.readback()
            dt = time.time() - t0
            st = ctx.stats()
            from sol_rs_b200 import io
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            io.write_png(os.path.join(ROOT, "gpurun_out", "tunnel_1080p_4096spp.png"), img)
            print(json.dumps({"config": "5-pathtrace tunnel.gltf --sky 1920x1080, 512 frames = 4096 spp, cap 8 (configs[2] in full)",
                              "seconds": dt, "ms_per_frame": 1e3 * dt / 512, "Mrays_s": st.rays / dt / 1e6, "rays": int(st.rays),
                              "mean_rgb": [float(v) for v in accum.readback()[..., :3].mean(axis=(0, 1))]}), flush=True)
        elif what in ("mega_cornell", "mega_tunnel", "wf_cornell"):
            # 1080p cap 8 with the schedule forced (SOLB_MEGA_VOTE / SOLB_MEGA_PERSISTENT select the megakernel variant)
            name = what.split("_")[1]
            sc = scene.load_scene(ctx, os.path.join(ROOT, "assets/models/%s.gltf" % name))
            sd = ray.SceneDescription.from_scene(ctx, sc)
            sched = N.SCHEDULE_WAVEFRONT if what.startswith("wf") else N.SCHEDULE_MEGAKERNEL
            r = pathtrace(sd, sc.camera, 1920, 1080, name == "tunnel", 8, args.frames, sched)
            print(json.dumps({"config": "5-pathtrace %s.gltf 1920x1080 cap 8, %s, SOLB_MEGA_VOTE=%s SOLB_MEGA_PERSISTENT=%s" % (
                name, "wavefront" if what.startswith("wf") else "megakernel", os.environ.get("SOLB_MEGA_VOTE", "default"),
                os.environ.get("SOLB_MEGA_PERSISTENT", "default")), **r}), flush=True)
        elif what == "cornell512":
            sc = scene.load_scene(ctx, os.path.join(ROOT, "assets/models/cornell.gltf"))
            sd = ray.SceneDescription.from_scene(ctx, sc)
            r = pathtrace(sd, sc.camera, 512, 512, False, 4, 8)
            print(json.dumps({"config": "5-pathtrace cornell.gltf 512x512 cap 4 (configs[0] on the GPU)", **r}), flush=True)
        elif what == "synth":
            t0 = time.time()
            sc = synth.make_scene(args.blas, args.grid)
            gen_s = time.time() - t0
            t0 = time.time()
            sd = ray.SceneDescription.from_scene(ctx, sc)
            ctx.synchronize()
            create_s = time.time() - t0
            info = sd.accel_info()
            st = ctx.stats()
            r = pathtrace(sd, sc.camera, 1920, 1080, True, 8, args.frames)
            print(json.dumps({"config": "synthetic %d BLAS x %d tris = %d triangles, 1080p --sky cap 8" % (args.blas, 2 * args.grid ** 2, info.n_triangles),
                              "build_ms": st.last_build_ms, "Mtris_s": info.n_triangles / (st.last_build_ms * 1e-3) / 1e6,
                              "n_wide_nodes": info.n_wide_nodes, "wide_depth": info.wide_depth, "sah_lbvh": info.sah_cost_lbvh,
                              "sah": info.sah_cost_binary, "scene_gen_s": gen_s, "upload_plus_build_s": create_s, **r}), flush=True)
            del sd, sc
        elif what == "build":
            for name in ("cornell", "Duck", "tunnel"):
                sc = scene.load_scene(ctx, os.path.join(ROOT, "assets/models/%s.gltf" % name))
                sd = ray.SceneDescription.from_scene(ctx, sc)
                best = 1e9
                for _ in range(5):
                    sd.blas_transform(sc.meshes[0].transform, 0)  # marks dirty -> full rebuild
                    sd.tlas_regenerate()
                    best = min(best, ctx.stats().last_build_ms)
                info = sd.accel_info()
                print(json.dumps({"config": "build " + name, "triangles": info.n_triangles, "build_ms": best,
                                  "Mtris_s": info.n_triangles / (best * 1e-3) / 1e6, "sah_lbvh": info.sah_cost_lbvh,
                                  "sah": info.sah_cost_binary, "nodes": info.n_wide_nodes}), flush=True)


if __name__ == "__main__":
    main()
