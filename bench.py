#!/usr/bin/env python
"""bench.py — headline benchmark of the sol-rs ray-tracing hot path on B200.

Metric (BASELINE.json): Mrays/s and ms/frame, 1080p 5-pathtrace.  A *step* is one reference frame: one
cmd_trace_rays launch sequence over W x H pixels x 8 samples (assets/glsl/pathtrace.rgen:39-104) of
tunnel.gltf --sky at 1920x1080 with max_bounces = 8 (BASELINE.json configs[2]); a *ray* is one
traceRayEXT-equivalent (traversal + its hit/miss shading).  cornell.gltf at the same size is reported
beside it in "extra".  One JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--schedule wavefront|megakernel]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WIDTH, HEIGHT, SPP, MAX_BOUNCES = 1920, 1080, 8, 8
WORKLOAD = "5-pathtrace tunnel.gltf --sky %dx%d, %d spp/frame, max_bounces %d" % (WIDTH, HEIGHT, SPP, MAX_BOUNCES)
CPU_SAMPLE_W, CPU_SAMPLE_H = 480, 270  # bounded sample of the same workload for the CPU arms


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(steps, warmup, sample_wh):
    """The reference's path on host cores.  The reference itself cannot run here (no rustc / Vulkan / lavapipe),
    so this is the oracle port (oracle/oracle.c, OpenMP on all host threads) on a bounded sample of the workload:
    the same scene / camera / seeds / spp / bounces at a reduced resolution."""
    import oracle
    from oracle import camera as ocam
    from oracle import gltf_flatten as gf

    w, h = sample_wh
    # all the host threads this process may use, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    oracle.lib().orc_set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    fs = gf.load_scene(os.path.join(ROOT, "assets", "models", "tunnel.gltf"))
    sc = oracle.Scene(fs)
    cam = ocam.Camera.from_view(fs.camera["view"], fs.camera["yfov"], fs.camera["znear"], fs.camera["zfar"])
    cam.set_window_size((w, h))
    acc = np.zeros((h, w, 4), dtype=np.float32)
    for f in range(warmup):
        sc.pathtrace_frame(ocam.scene_uniforms(cam, w, h, f), w, h, acc, 0, True, SPP, MAX_BOUNCES)
    st = oracle.OrcStats()
    t0 = time.perf_counter()
    for f in range(steps):
        sc.pathtrace_frame(ocam.scene_uniforms(cam, w, h, warmup + f), w, h, acc, 0, True, SPP, MAX_BOUNCES, st)
    dt = time.perf_counter() - t0
    return {"value": st.rays / dt / 1e6, "unit": "Mrays/s", "cores": oracle.lib().orc_num_threads(), "kind": "port",
            "sample": "%dx%d of the 1920x1080 frame (same scene/camera/seeds/spp/bounces), %d frame(s), %.1f s" % (w, h, steps, dt),
            "ms_per_step": 1e3 * dt / max(steps, 1), "rays": int(st.rays)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))  # bounded: ~1.6 s per sample frame on 16 host threads, ~6 s on 8
    cb = cpu_reference_arm(steps, min(args.warmup, 1), (CPU_SAMPLE_W, CPU_SAMPLE_H))
    line = {"impl": "reference", "metric": "Mrays/s", "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference GLSL/driver path cannot run here (no rustc/Vulkan/lavapipe); "
                       "this arm times the CPU oracle port on a bounded sample"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


_JSON_FD = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries write there too (NCCL prints its version banner with
    printf).  Keep a private duplicate of fd 1 for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="solb")
    ap.add_argument("--schedule", default=os.environ.get("SOLB_SCHEDULE", "wavefront"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also bench cornell and the other schedule")
    ap.add_argument("--workload", default="tunnel", choices=["tunnel", "synth"],
                    help="tunnel = the headline config; synth = BASELINE configs[4]: synthetic instanced scene (--blas x 20000 triangles)")
    ap.add_argument("--blas", type=int, default=1000)
    ap.add_argument("--split", default="frames", choices=["frames", "tiles"],
                    help="multi-GPU split: frames = rank r renders frames f = r (mod N), one NCCL reduce at the end (default); tiles = "
                         "every frame is cut into N row tiles, one all-gather of the accumulation rows per frame (single-sample mode)")
    ap.add_argument("--accel", default="flat", choices=["flat", "two_level"],
                    help="flat = transforms baked into one hierarchy (default); two_level = TLAS over object-space BLASes")
    ap.add_argument("--size", default="", help="WxH override, e.g. 3840x2160 for BASELINE configs[3] (default 1920x1080)")
    args = ap.parse_args()
    global WIDTH, HEIGHT, WORKLOAD
    if args.size:
        WIDTH, HEIGHT = (int(v) for v in args.size.lower().split("x"))
        WORKLOAD = "5-pathtrace tunnel.gltf --sky %dx%d, %d spp/frame, max_bounces %d" % (WIDTH, HEIGHT, SPP, MAX_BOUNCES)
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import sol_rs_b200 as sol
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import multigpu, ray, scene

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsolb has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # NCCL prints its version banner (NCCL_DEBUG >= VERSION) on stdout, which must carry exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL logs to stdout by default
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # a real (non-default) torch stream, made current: libsolb launches on it, torch events time it, NCCL orders with it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = sol.Context(local_rank, stream.cuda_stream)
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    sched = {"mega": N.SCHEDULE_MEGAKERNEL, "wave": N.SCHEDULE_WAVEFRONT, "warp": N.SCHEDULE_WARPFRONT}[args.schedule[:4]]
    sched_name = {N.SCHEDULE_MEGAKERNEL: "megakernel", N.SCHEDULE_WAVEFRONT: "wavefront", N.SCHEDULE_WARPFRONT: "warpfront"}[sched]
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def setup(model, sky):
        if model == "synth":
            from sol_rs_b200 import synth

            sc = synth.make_scene(args.blas, 100)
        else:
            sc = scene.load_scene(ctx, os.path.join(ROOT, "assets", "models", model))
        sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=N.ACCEL_TWO_LEVEL if args.accel == "two_level" else N.ACCEL_FLAT)
        cam = sc.camera
        cam.set_window_size((WIDTH, HEIGHT))
        pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/pathtrace.rgen", ray.RAYGEN_KHR)
                            .shader("glsl/pathtrace.rmiss", ray.MISS_KHR).shader("glsl/pathtrace.rchit", ray.CLOSEST_HIT_KHR)
                            .specialization([1 if sky else 0], 0))
        sbt = ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))
        return sc, sd, cam, sbt

    def device_timed(sd, cam, sbt, schedule, n_warm, n_steps, mode, frame0):
        """K steps with inputs resident in HBM; per-step CUDA events on the launching stream; L2 flushed between steps."""
        accum = sol.Image2d(ctx, WIDTH, HEIGHT, N.FORMAT_RGBA32F)
        render = sol.Image2d(ctx, WIDTH, HEIGHT, N.FORMAT_RGBA8)
        tiled = dist is not None and args.split == "tiles"
        if tiled:
            return device_timed_tiles(sd, cam, sbt, schedule, n_warm, n_steps, frame0, accum, render)
        frames = multigpu.frames_for_rank(rank, world, world * (n_warm + n_steps), first=frame0)  # f = r (mod R): SURVEY 8e
        for f in frames[:n_warm]:
            sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, WIDTH, HEIGHT, f), accum, render, schedule=schedule,
                                                 samples_per_frame=SPP, max_bounces=MAX_BOUNCES, accum_mode=mode), (WIDTH, HEIGHT, 1))
        accum.clear()
        ctx.reset_stats()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps + 1)]
        t_wall = time.perf_counter()
        for i, f in enumerate(frames[n_warm:]):
            l2_flush.fill_(i & 0xFF)  # > 126 MB L2, outside the per-step event pair
            u = scene.scene_uniforms(cam, WIDTH, HEIGHT, f)
            evs[i][0].record(stream)
            sd.tlas_regenerate()  # reference rebuilds the TLAS every frame (examples/5-pathtrace.rs:316); no-op when clean
            sbt.cmd_trace_rays(ray.TraceBindings(sd, u, accum, render, schedule=schedule, samples_per_frame=SPP,
                                                 max_bounces=MAX_BOUNCES, accum_mode=mode), (WIDTH, HEIGHT, 1))
            evs[i][1].record(stream)
        reduce_ms = 0.0
        if dist:
            # the one real exchange of the path: sum the per-rank accumulation buffers over NVLink, resolve on rank 0
            evs[n_steps][0].record(stream)
            multigpu.reduce_accum(accum.as_torch(), dst=0)
            if rank == 0:
                ray.resolve_sum(ctx, accum, accum, render)
            evs[n_steps][1].record(stream)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_wall
        step_ms = [a.elapsed_time(b) for a, b in evs[:n_steps]]
        if dist:
            reduce_ms = evs[n_steps][0].elapsed_time(evs[n_steps][1])
        st = ctx.stats()
        return {"ms_total": sum(step_ms) + reduce_ms, "step_ms": step_ms, "reduce_ms": reduce_ms, "rays": int(st.rays),
                "paths": int(st.paths), "hits": int(st.hits), "launches": int(st.kernel_launches), "wall_s": wall}

    def device_timed_tiles(sd, cam, sbt, schedule, n_warm, n_steps, frame0, accum, render):
        """Tile split (single-sample interactive mode): every rank traces rows [r H/N, (r+1) H/N) of EVERY frame into
        the full-size targets, then one all-gather of the accumulation rows makes the frame whole on every rank."""
        band = 8  # rows per band: rank r owns bands r, r + N, r + 2N, ... (interleaved: the image's cheap and expensive rows are
        #           shared out evenly; contiguous tiles measured 10.6 ms at 8 GPUs)
        full = accum.as_torch()
        n_bands = (HEIGHT + band - 1) // band
        per_rank = (n_bands + world - 1) // world
        rows_of = []
        for r in range(world):
            rows = [min(b * band + j, HEIGHT - 1) for b in range(r, n_bands, world) for j in range(band)]
            rows += [rows[-1]] * (per_rank * band - len(rows))  # pad to a common size (duplicates rewrite the same row)
            rows_of.append(torch.tensor(rows, dtype=torch.long, device="cuda"))
        all_rows = torch.cat(rows_of)
        gathered = torch.empty((world, per_rank * band, WIDTH, 4), dtype=torch.float32, device="cuda")
        parts = list(gathered.unbind(0))

        def one(f):
            u = scene.scene_uniforms(cam, WIDTH, HEIGHT, f)
            sbt.cmd_trace_rays(ray.TraceBindings(sd, u, accum, render, schedule=schedule, samples_per_frame=SPP, max_bounces=MAX_BOUNCES,
                                                 tile_rows=(rank * band, band, world * band)), (WIDTH, HEIGHT, 1))
            mine = full.index_select(0, rows_of[rank])       # this rank's bands, compacted
            dist.all_gather(parts, mine)
            full.index_copy_(0, all_rows, gathered.view(-1, WIDTH, 4))  # every rank now holds the whole frame

        for f in range(frame0, frame0 + n_warm):
            one(f)
        accum.clear()
        ctx.reset_stats()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        t_wall = time.perf_counter()
        for i in range(n_steps):
            l2_flush.fill_(i & 0xFF)
            evs[i][0].record(stream)
            one(frame0 + n_warm + i)
            evs[i][1].record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_wall
        step_ms = [a.elapsed_time(b) for a, b in evs]
        st = ctx.stats()
        return {"ms_total": sum(step_ms), "step_ms": step_ms, "reduce_ms": 0.0, "rays": int(st.rays), "paths": int(st.paths),
                "hits": int(st.hits), "launches": int(st.kernel_launches), "wall_s": wall}

    def gather_max_sum(ms_total, rays):
        if not dist:
            return ms_total, rays
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        r = torch.tensor([rays], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        return float(t.item()), int(r.item())

    sc, sd, cam, sbt = setup("synth" if args.workload == "synth" else "tunnel.gltf", True)
    build_ms = ctx.stats().last_build_ms
    workload = WORKLOAD if args.workload == "tunnel" else (
        "5-pathtrace synthetic %d BLAS x 20000 triangles (seed 0xB200) --sky %dx%d, %d spp/frame, max_bounces %d" % (
            args.blas, WIDTH, HEIGHT, SPP, MAX_BOUNCES))
    mode = N.ACCUM_SUM if world > 1 else N.ACCUM_MIX

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    main_run = device_timed(sd, cam, sbt, sched, warmup, steps, mode, 0)
    clock_info = clocks.stop() if rank == 0 else None
    ms_total, rays_total = gather_max_sum(main_run["ms_total"], main_run["rays"])
    value = rays_total / (ms_total * 1e-3) / 1e6
    ms_per_step = ms_total / steps

    # ---- e2e (every rank): the call a user makes, host buffers, H2D of the step's inputs + D2H of the rendered frame each
    #      step; whole-job value = rays of all ranks / the slowest rank's wall clock ----
    e2e_accum = sol.Image2d(ctx, WIDTH, HEIGHT, N.FORMAT_RGBA32F)
    e2e_render = sol.Image2d(ctx, WIDTH, HEIGHT, N.FORMAT_RGBA8)
    host_frame = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8, pin_memory=True).numpy()
    for f in range(2):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, WIDTH, HEIGHT, f), e2e_accum, e2e_render, schedule=sched,
                                             samples_per_frame=SPP, max_bounces=MAX_BOUNCES), (WIDTH, HEIGHT, 1))
    ctx.reset_stats()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    n_e2e = min(steps, 8)
    for f in multigpu.frames_for_rank(rank, world, world * n_e2e, first=200):
        u = scene.scene_uniforms(cam, WIDTH, HEIGHT, f)  # host-side camera -> 400-byte uniform block (the step's input)
        sd.tlas_regenerate()
        sbt.cmd_trace_rays(ray.TraceBindings(sd, u, e2e_accum, e2e_render, schedule=sched, samples_per_frame=SPP,
                                             max_bounces=MAX_BOUNCES), (WIDTH, HEIGHT, 1))
        e2e_render.readback(host_frame)  # replaces blit-to-present: the step's result reaches host memory
    dt_e2e = time.perf_counter() - t0
    e2e_ms, e2e_rays = gather_max_sum(1e3 * dt_e2e, int(ctx.stats().rays))
    e2e = {"value": e2e_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 400 + 32,
           "d2h_bytes_per_step": WIDTH * HEIGHT * 4, "ms_per_step": e2e_ms / n_e2e, "steps": n_e2e}
    del e2e_accum, e2e_render

    # ---- roofline of the dominant kernel: instrumented pass for nodes/triangles per ray + per-launch event timing ----
    roofline = None
    cpu_baseline = None
    extra = {}
    if rank == 0:
        ctx.reset_stats()
        accum = sol.Image2d(ctx, WIDTH, HEIGHT, N.FORMAT_RGBA32F)
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, WIDTH, HEIGHT, 0), accum, None, schedule=sched,
                                             samples_per_frame=SPP, max_bounces=MAX_BOUNCES, collect_stats=True), (WIDTH, HEIGHT, 1))
        st = ctx.stats()
        n_node, n_tri = st.nodes_visited / st.rays, st.tris_tested / st.rays
        p_hit, r_path = st.hits / st.rays, st.rays / st.paths
        # live per-launch timing of the dominant kernel (event pairs inside libsolb, timing mode)
        ctx.set_timing(True)
        ctx.reset_stats()
        for f in range(100, 103):
            sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, WIDTH, HEIGHT, f), accum, None, schedule=sched,
                                                 samples_per_frame=SPP, max_bounces=MAX_BOUNCES), (WIDTH, HEIGHT, 1))
        st2 = ctx.stats()
        ctx.set_timing(False)
        if sched in (N.SCHEDULE_MEGAKERNEL, N.SCHEDULE_WARPFRONT):
            kernel = "k_pathtrace_mega" if sched == N.SCHEDULE_MEGAKERNEL else "k_pt_warpfront"
            # everything happens in one kernel: traversal + shading fetches + accumulation RMW (+ the path record round trip
            # of the warp-local wavefront: ray record read, hit record write / read, path state read / write)
            b_ray = 80 * n_node + 48 * n_tri + p_hit * (112 + 176) + 36.0 / (SPP * r_path) + (
                (32 + 32 + 80 + 64) if sched == N.SCHEDULE_WARPFRONT else 0)
        else:
            kernel = "k_wf_trace"
            # traversal kernel only: nodes + triangles + ray record read (2 x float4 + queue id) + hit record write
            b_ray = 80 * n_node + 48 * n_tri + 36 + 16
        rays_per_launch = st2.rays / max(st2.trace_kernel_launches, 1)
        avg_launch_ms = st2.trace_kernel_ms_total / max(st2.trace_kernel_launches, 1)
        achieved = b_ray * rays_per_launch / (avg_launch_ms * 1e-3) / 1e9
        peak, peak_src = measured_peak_gbs()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f)[kernel]["dram_bytes_per_launch"]  # from the committed ncu --set full capture
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "algorithmic_bytes_per_launch": b_ray * rays_per_launch, "peak_source": peak_src, "bytes_per_ray": b_ray, "nodes_per_ray": n_node, "tris_per_ray": n_tri,
                    "p_hit": p_hit, "rays_per_path": r_path, "avg_launch_ms": avg_launch_ms,
                    "launches_timed": int(st2.trace_kernel_launches),
                    "kernel_share_of_step": st2.trace_kernel_ms_total / max(sum(main_run["step_ms"][:3]), 1e-9) if steps >= 3 else None,
                    "note": "algorithmic bytes are served mostly from L1/L2 (BVH + triangles = %.1f MB): frac can exceed what DRAM counters show" % (
                        (sd.accel_info().n_wide_nodes * 80 + sd.accel_info().n_triangles * 48) / 1e6)}

        if world == 1 and args.workload == "tunnel" and not args.size:
            # the metric names cornell beside tunnel: same resolution / spp / bounce cap, default (auto) schedule
            _, sd_c, cam_c, sbt_c = setup("cornell.gltf", False)
            rc = device_timed(sd_c, cam_c, sbt_c, N.SCHEDULE_AUTO, 3, min(steps, 8), N.ACCUM_MIX, 0)
            extra["cornell_1080p_cap8"] = {"Mrays_s": rc["rays"] / (rc["ms_total"] * 1e-3) / 1e6, "ms_per_frame": rc["ms_total"] / min(steps, 8),
                                           "rays_per_path": rc["rays"] / max(rc["paths"], 1), "schedule": "auto (megakernel: 3 wide nodes)"}
        if args.extra:
            other = N.SCHEDULE_WAVEFRONT if sched == N.SCHEDULE_MEGAKERNEL else N.SCHEDULE_MEGAKERNEL
            r2 = device_timed(sd, cam, sbt, other, 2, min(steps, 6), N.ACCUM_MIX, 0) if world == 1 else None
            if r2:
                extra["other_schedule"] = {"schedule": "megakernel" if other else "wavefront",
                                           "Mrays_s": r2["rays"] / (r2["ms_total"] * 1e-3) / 1e6,
                                           "ms_per_step": r2["ms_total"] / min(steps, 6)}
            if world == 1:
                _, sd_c, cam_c, sbt_c = setup("cornell.gltf", False)
                for name, s_ in (("wavefront", N.SCHEDULE_WAVEFRONT), ("megakernel", N.SCHEDULE_MEGAKERNEL)):
                    rc = device_timed(sd_c, cam_c, sbt_c, s_, 2, min(steps, 6), N.ACCUM_MIX, 0)
                    extra["cornell_1080p_b8_" + name] = {"Mrays_s": rc["rays"] / (rc["ms_total"] * 1e-3) / 1e6,
                                                         "ms_per_step": rc["ms_total"] / min(steps, 6),
                                                         "rays_per_path": rc["rays"] / max(rc["paths"], 1)}
        if not args.no_cpu_baseline and world == 1 and args.workload == "tunnel":
            cb = cpu_reference_arm(6, 1, (CPU_SAMPLE_W, CPU_SAMPLE_H))  # ~10 s on 16 host threads
            cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        tiled = world > 1 and args.split == "tiles"
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if tiled else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "schedule": sched_name, "accel": args.accel, "bvh_build_ms": build_ms,
                           "frames_per_rank": steps, "rays_per_frame": rays_total / (steps * (1 if tiled else world)),
                           "l2": "256 MB buffer written between timed steps (outside the per-step event pairs)",
                           "multi_gpu": ("every frame cut into 8-row bands dealt round-robin to the N ranks, one NCCL all-gather of the accumulation rows per frame inside the timed region"
                                         if tiled else
                                         "frames f = rank (mod N) per rank, local sums, one NCCL reduce + resolve inside the timed region")
                           if world > 1 else "single GPU"},
                "ms_per_frame": ms_per_step, "reduce_ms": main_run["reduce_ms"],
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": main_run["launches"],
                "clocks": clock_info, "extra": extra}
        _emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
