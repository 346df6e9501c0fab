"""Time one rank's share of a tile-split frame on a single GPU (tunnel 1080p, interleaved 8-row bands of an N-way split).
    SOLB_OVERLAP=1 python tools/tile_time.py 8"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sol_rs_b200 as sol  # noqa: E402
from sol_rs_b200 import _native as N, ray, scene  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SCHED = {"wavefront": N.SCHEDULE_WAVEFRONT, "warpfront": N.SCHEDULE_WARPFRONT, "megakernel": N.SCHEDULE_MEGAKERNEL}[sys.argv[2] if len(sys.argv) > 2 else "warpfront"]
W, H = 1920, 1080
ctx = sol.Context(0)
sc = scene.load_scene(ctx, os.path.join(ROOT, "assets", "models", "tunnel.gltf"))
sd = ray.SceneDescription.from_scene(ctx, sc)
cam = sc.camera
cam.set_window_size((W, H))
pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/pathtrace.rgen", ray.RAYGEN_KHR).shader("glsl/pathtrace.rmiss", ray.MISS_KHR)
                    .shader("glsl/pathtrace.rchit", ray.CLOSEST_HIT_KHR).specialization([1], 0))
sbt = ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))
accum, render = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F), sol.Image2d(ctx, W, H, N.FORMAT_RGBA8)
for rank in (0, world // 2):
    tile = (rank * 8, 8, world * 8) if world > 1 else None
    for f in range(3):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, W, H, f), accum, render, max_bounces=8, schedule=SCHED,
                                             tile_rows=tile), (W, H, 1))
    ctx.synchronize()
    ctx.reset_stats()
    t0 = time.perf_counter()
    n = 10
    for f in range(3, 3 + n):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, W, H, f), accum, render, max_bounces=8, schedule=SCHED,
                                             tile_rows=tile), (W, H, 1))
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("world %d rank %d: %.2f ms per tile-frame, %.0f Mrays/s on this GPU" % (world, rank, 1e3 * dt, ctx.stats().rays / n / dt / 1e6))
