"""Where a tile-split frame's time goes at N GPUs (torchrun): one rank's share traced alone, + the per-frame all-gather of the
rgba8 frame, + that of the float accumulation, with and without the L2 flush between frames.  Max over ranks, ms per frame.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/tile_split_bench.py [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sol_rs_b200 as sol  # noqa: E402
from sol_rs_b200 import _native as N, multigpu, ray, scene  # noqa: E402

W, H, BAND = 1920, 1080, 8
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 48
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = sol.Context(local, stream.cuda_stream)
ctx.preload()
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = multigpu.Communicator.from_torch_distributed(ctx)
sc = scene.load_scene(ctx, os.path.join(ROOT, "assets", "models", "tunnel.gltf"))
sd = ray.SceneDescription.from_scene(ctx, sc)
cam = sc.camera
cam.set_window_size((W, H))
pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/pathtrace.rgen", ray.RAYGEN_KHR).shader("glsl/pathtrace.rmiss", ray.MISS_KHR)
                    .shader("glsl/pathtrace.rchit", ray.CLOSEST_HIT_KHR).specialization([1], 0))
sbt = ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))
accum, render = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F), sol.Image2d(ctx, W, H, N.FORMAT_RGBA8)
tile = multigpu.tile_rows_for_rank(rank, world, BAND)
l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(gather, flush, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for f in range(n):
        if flush:
            l2_flush.fill_(f & 0xFF)
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, W, H, f), accum, render, max_bounces=8, tile_rows=tile), (W, H, 1))
        if gather == "render":
            comm.allgather_rows(render, BAND)
        elif gather == "accum":
            comm.allgather_rows(accum, BAND)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


run("render", True, 4)
for gather in ("none", "render", "accum"):
    for flush in (False, True):
        ms = run(gather, flush, frames)
        if rank == 0:
            print("world %d gather=%-6s l2_flush=%d : %.3f ms per tile-frame" % (world, gather, flush, ms), flush=True)
comm.close()
dist.destroy_process_group()
