#!/usr/bin/env python
"""bench.py — headline benchmark of the sol-rs ray-tracing hot path on B200.

Metric (BASELINE.json): Mrays/s and ms/frame, 1080p 5-pathtrace.  A *step* is one reference frame: one
cmd_trace_rays launch sequence over W x H pixels x 8 samples (assets/glsl/pathtrace.rgen:39-104) of
tunnel.gltf --sky at 1920x1080 with max_bounces = 8 (BASELINE.json configs[2]); a *ray* is one
traceRayEXT-equivalent (traversal + its hit/miss shading).  cornell.gltf at the same size is reported
beside it in "extra".  One JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--schedule auto|warpfront|wavefront|megakernel]

The JSON line's `extra` holds, at N = 1, the other BASELINE.json configs each with its own roofline (cornell 1080p, configs[1]
4-ray-ao, configs[3] 4K, configs[4] synthetic 20 M triangles incl. its BVH build) and, at N > 1, the tile split of single frames
(`extra.tile_split`).  The multi-GPU exchange runs through the C ABI (solb_reduce_accum / solb_allgather_rows, NCCL underneath);
torch.distributed only bootstraps the communicator id, provides the barriers and gathers the per-rank timings.
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
    steps = max(1, min(args.steps, 64))  # one step = one frame of the bounded sample: ~1.6 s on 16 host threads
    cb = cpu_reference_arm(steps, min(args.warmup, 1), (CPU_SAMPLE_W, CPU_SAMPLE_H))
    line = {"impl": "reference", "metric": "Mrays/s", "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "%dx%d pixels of the %dx%d frame per step" % (CPU_SAMPLE_W, CPU_SAMPLE_H, WIDTH, HEIGHT),
                       "note": "reference GLSL/driver path cannot run here (no rustc/Vulkan/lavapipe); "
                       "this arm times the CPU oracle port (oracle/oracle.c, OpenMP) on a bounded sample of every frame"},
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="solb")
    ap.add_argument("--schedule", default=os.environ.get("SOLB_SCHEDULE", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs in `extra`")
    ap.add_argument("--extra", action="store_true", help="also bench the other schedules on tunnel and cornell")
    ap.add_argument("--workload", default="tunnel", choices=["tunnel", "synth"],
                    help="tunnel = the headline config; synth = BASELINE configs[4]: synthetic instanced scene (--blas x 20000 triangles)")
    ap.add_argument("--blas", type=int, default=1000)
    ap.add_argument("--split", default="frames", choices=["frames", "tiles"],
                    help="multi-GPU split of the headline value: frames = rank r renders frames f = r (mod N), one NCCL reduce at the end "
                         "(default); tiles = every frame is cut into N sets of row bands, one all-gather per frame")
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
    from sol_rs_b200 import io as sol_io
    from sol_rs_b200 import multigpu, ray, scene

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsolb has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    # a real (non-default) torch stream, made current: libsolb launches on it, torch events time it, NCCL runs on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = sol.Context(local_rank, stream.cuda_stream)
    ctx.preload()  # load the kernels now (what device / pipeline creation is to the reference), not inside the first build
    comm = None
    if world > 1:
        # NCCL prints its version banner (NCCL_DEBUG >= VERSION) on stdout, which must carry exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = multigpu.Communicator.from_torch_distributed(ctx)  # the path's exchange: libsolb's own communicator (C ABI)
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    sched = {"mega": N.SCHEDULE_MEGAKERNEL, "wave": N.SCHEDULE_WAVEFRONT, "warp": N.SCHEDULE_WARPFRONT, "auto": N.SCHEDULE_AUTO}[args.schedule[:4]]
    sched_names = {N.SCHEDULE_MEGAKERNEL: "megakernel", N.SCHEDULE_WAVEFRONT: "wavefront", N.SCHEDULE_WARPFRONT: "warpfront",
                   N.SCHEDULE_AUTO: "auto"}
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    BAND = 8  # tile split: rows per band; rank r owns bands r, r + N, ... (interleaved: cheap and expensive rows are shared out)

    def pt_sbt(sky):
        pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/pathtrace.rgen", ray.RAYGEN_KHR)
                            .shader("glsl/pathtrace.rmiss", ray.MISS_KHR).shader("glsl/pathtrace.rchit", ray.CLOSEST_HIT_KHR)
                            .specialization([1 if sky else 0], 0))
        return ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))

    def setup(model, sky, accel=None):
        if model == "synth":
            from sol_rs_b200 import synth

            sc = synth.make_scene(args.blas, 100)
        else:
            sc = scene.load_scene(ctx, os.path.join(ROOT, "assets", "models", model))
        two = (accel or args.accel) == "two_level"
        sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=N.ACCEL_TWO_LEVEL if two else N.ACCEL_FLAT)
        return sc, sd, sc.camera, pt_sbt(sky)

    def resolved_schedule(sd, schedule):
        if schedule != N.SCHEDULE_AUTO:
            return schedule
        auto_wide = N.SCHEDULE_WAVEFRONT if os.environ.get("SOLB_AUTO_SCHEDULE", "3") == "0" else N.SCHEDULE_WARPFRONT
        return N.SCHEDULE_MEGAKERNEL if sd.accel_info().n_wide_nodes <= 8 else auto_wide

    def barrier_sync():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def device_timed(sd, cam, sbt, schedule, n_warm, n_steps, split, frame0, w=None, h=None, mb=None, trace_fn=None):
        """K steps with inputs resident in HBM; CUDA events on the launching stream; L2 flushed between steps.  The clock is
        ONE interval from before the first step to after the last (steps overlap when the schedule keeps frames in flight, so
        per-step intervals would leave the work done during the flushes uncounted); the flushes are inside it.
        split: None (this GPU alone), "frames" (f = r mod N + one reduce + resolve at the end) or "tiles" (bands + all-gather)."""
        w, h, mb = w or WIDTH, h or HEIGHT, MAX_BOUNCES if mb is None else mb
        cam.set_window_size((w, h))
        accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
        render = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
        n_ranks = world if split else 1
        mode = N.ACCUM_SUM if split == "frames" and world > 1 else N.ACCUM_MIX
        tile = multigpu.tile_rows_for_rank(rank, world, BAND) if split == "tiles" and world > 1 else None
        if split == "frames":
            frames = multigpu.frames_for_rank(rank, n_ranks, n_ranks * (n_warm + n_steps), first=frame0)  # f = r (mod R): SURVEY 8e
        else:
            frames = list(range(frame0, frame0 + n_warm + n_steps))

        def one(f):
            sd.tlas_regenerate()  # reference rebuilds the TLAS every frame (examples/5-pathtrace.rs:316); no-op when clean
            if trace_fn:
                trace_fn(f, accum)
            else:
                sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, render, schedule=schedule,
                                                     samples_per_frame=SPP, max_bounces=mb, accum_mode=mode, tile_rows=tile), (w, h, 1))
            if tile:
                comm.allgather_rows(render, BAND)  # every rank now holds the whole displayable frame (what the reference presents);
                # the float accumulation stays distributed, every rank owning its bands, and is gathered once at the end

        for f in frames[:n_warm]:
            one(f)
        accum.clear()
        ctx.reset_stats()
        barrier_sync()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps + 1)]
        t_wall = time.perf_counter()
        for i, f in enumerate(frames[n_warm:]):
            l2_flush.fill_(i & 0xFF)  # > 126 MB L2
            evs[i][0].record(stream)
            one(f)
            evs[i][1].record(stream)
        reduce_ms = 0.0
        if tile:
            comm.allgather_rows(accum, BAND)
            evs[n_steps - 1][1].record(stream)  # (re-recorded: the closing gather of the accumulation is inside the clock)
        if split == "frames" and world > 1:
            # the one real exchange of the path: sum the per-rank accumulation buffers over NVLink, resolve on rank 0
            evs[n_steps][0].record(stream)
            comm.reduce_accum(accum, 0, accum, render)
            evs[n_steps][1].record(stream)
        barrier_sync()
        wall = time.perf_counter() - t_wall
        step_ms = [a.elapsed_time(b) for a, b in evs[:n_steps]]
        last = evs[n_steps - 1][1]
        if split == "frames" and world > 1:
            reduce_ms = evs[n_steps][0].elapsed_time(evs[n_steps][1])
            last = evs[n_steps][1]
        st = ctx.stats()
        return {"ms_total": evs[0][0].elapsed_time(last), "step_ms": step_ms, "reduce_ms": reduce_ms, "rays": int(st.rays),
                "paths": int(st.paths), "hits": int(st.hits), "launches": int(st.kernel_launches), "wall_s": wall}

    def gather_max_sum(ms_total, rays):
        if not dist:
            return ms_total, rays
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        r = torch.tensor([rays], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        return float(t.item()), int(r.item())

    def load_traffic(kernel, workload_key):
        """dram bytes per launch from a committed `ncu --set full` capture of THIS kernel on THIS workload, else None"""
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                return json.load(f)[kernel][workload_key]["dram_bytes_per_launch"]
        except Exception:
            return None

    def measure_roofline(sd, cam, sbt, schedule, w, h, mb, workload_key, timed_ms_per_frame=None):
        """Instrumented pass (nodes / triangles per ray) + live per-launch timing of the dominant kernel (event pairs inside
        libsolb, on the ctx stream).  Algorithmic bytes per ray: SURVEY 8d's formula with this run's counters."""
        cam.set_window_size((w, h))
        ctx.reset_stats()
        accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, 0), accum, None, schedule=schedule,
                                             samples_per_frame=SPP, max_bounces=mb, collect_stats=True), (w, h, 1))
        st = ctx.stats()
        n_node, n_tri = st.nodes_visited / st.rays, st.tris_tested / st.rays
        p_hit, r_path = st.hits / st.rays, st.rays / st.paths
        ctx.set_timing(True)
        ctx.reset_stats()
        for f in range(100, 103):
            l2_flush.fill_(f & 0xFF)
            sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, None, schedule=schedule,
                                                 samples_per_frame=SPP, max_bounces=mb), (w, h, 1))
        st2 = ctx.stats()
        ctx.set_timing(False)
        real = resolved_schedule(sd, schedule)
        trace_only = 80 * n_node + 48 * n_tri + 36 + 16  # node + triangle fetches, ray record read (+ queue id), hit record write
        whole_ray = 80 * n_node + 48 * n_tri + 96 + 32 + p_hit * 204 + 32.0 / (SPP * r_path)  # SURVEY 8d, whole step
        if real == N.SCHEDULE_WAVEFRONT:
            kernel, b_ray, scope = "k_wf_trace", trace_only, "traversal kernel only (80 n + 48 t + 36 + 16)"
        else:
            kernel = "k_pathtrace_mega" if real == N.SCHEDULE_MEGAKERNEL else "k_pt_warpfront"
            b_ray, scope = whole_ray, "whole step in one kernel: SURVEY 8d formula 80 n + 48 t + 96 + 32 + 204 p_hit + 32 / (8 r)"
        rays_per_launch = st2.rays / max(st2.trace_kernel_launches, 1)
        avg_launch_ms = st2.trace_kernel_ms_total / max(st2.trace_kernel_launches, 1)
        achieved = b_ray * rays_per_launch / (avg_launch_ms * 1e-3) / 1e9
        peak, peak_src = measured_peak_gbs()
        info = sd.accel_info()
        return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": load_traffic(kernel, workload_key), "algorithmic_bytes_per_launch": b_ray * rays_per_launch,
                "peak_source": peak_src, "bytes_per_ray": b_ray, "bytes_per_ray_scope": scope, "nodes_per_ray": n_node, "tris_per_ray": n_tri,
                "p_hit": p_hit, "rays_per_path": r_path, "rays_per_launch": rays_per_launch, "avg_launch_ms": avg_launch_ms,
                "launches_timed": int(st2.trace_kernel_launches),
                "whole_step_bytes_per_ray": whole_ray,
                "bvh_MB": (info.n_wide_nodes * 80 + info.n_triangles * 48) / 1e6,
                "note": "algorithmic bytes; the BVH (%.1f MB) is served from L1/L2 when it fits, so frac can exceed what DRAM counters show"
                        % ((info.n_wide_nodes * 80 + info.n_triangles * 48) / 1e6)}

    sc, sd, cam, sbt = setup("synth" if args.workload == "synth" else "tunnel.gltf", True)
    build_ms = ctx.stats().last_build_ms
    rebuilds = []
    for _ in range(3):  # forced full rebuilds (what a per-frame TLAS::regenerate of a moving scene costs in flattened mode)
        sd.blas_transform(sc.meshes[0].transform, 0)
        sd.tlas_regenerate()
        rebuilds.append(ctx.stats().last_build_ms)
    workload = WORKLOAD if args.workload == "tunnel" else (
        "5-pathtrace synthetic %d BLAS x 20000 triangles (seed 0xB200) --sky %dx%d, %d spp/frame, max_bounces %d" % (
            args.blas, WIDTH, HEIGHT, SPP, MAX_BOUNCES))
    split = None if world == 1 else args.split
    tiled = split == "tiles"

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    main_run = device_timed(sd, cam, sbt, sched, warmup, steps, split, 0)
    clock_info = clocks.stop() if rank == 0 else None
    ms_total, rays_total = gather_max_sum(main_run["ms_total"], main_run["rays"])
    value = rays_total / (ms_total * 1e-3) / 1e6
    ms_per_step = ms_total / steps

    # ---- e2e (every rank): the call a user makes, host buffers: every step builds the 400-byte uniform block on the host
    #      (H2D with the launch) and reads the rendered frame back into pinned host memory (replaces blit-to-present); at
    #      N > 1 the closing reduce + resolve + readback of the combined image on rank 0 are inside the clock too.
    #      whole-job value = rays of all ranks / the slowest rank's wall clock ----
    cam.set_window_size((WIDTH, HEIGHT))
    e2e_accum = sol.Image2d(ctx, WIDTH, HEIGHT, N.FORMAT_RGBA32F)
    e2e_render = sol.Image2d(ctx, WIDTH, HEIGHT, N.FORMAT_RGBA8)
    # two frames in flight, like the reference's render loop (one fence + one present image per swapchain image,
    # src/renderer.rs:72-81, 188): frame f's read-back overlaps frame f+1's tracing, and a slot's host buffer is reused only
    # after its fence has been waited for
    host_frames = [ctx.host_alloc((HEIGHT, WIDTH, 4), np.uint8) for _ in range(2)]
    fences = [ctx.fence() for _ in range(2)]
    e2e_mode = N.ACCUM_SUM if world > 1 else N.ACCUM_MIX
    for f in range(2):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, WIDTH, HEIGHT, f), e2e_accum, e2e_render, schedule=sched,
                                             samples_per_frame=SPP, max_bounces=MAX_BOUNCES, accum_mode=e2e_mode), (WIDTH, HEIGHT, 1))
    e2e_accum.clear()
    ctx.reset_stats()
    barrier_sync()
    t0 = time.perf_counter()
    n_e2e = steps
    for i, f in enumerate(multigpu.frames_for_rank(rank, world, world * n_e2e, first=200)):
        fences[i & 1].wait()  # wait_for_and_reset_fence: the frame that last used this slot has reached host memory
        u = scene.scene_uniforms(cam, WIDTH, HEIGHT, f)  # host-side camera -> 400-byte uniform block (the step's input)
        sd.tlas_regenerate()
        sbt.cmd_trace_rays(ray.TraceBindings(sd, u, e2e_accum, e2e_render, schedule=sched, samples_per_frame=SPP,
                                             max_bounces=MAX_BOUNCES, accum_mode=e2e_mode), (WIDTH, HEIGHT, 1))
        e2e_render.readback_async(host_frames[i & 1])  # the step's result -> host memory (replaces blit + present)
        fences[i & 1].signal()
    if world > 1:
        comm.reduce_accum(e2e_accum, 0, e2e_accum, e2e_render)
        if rank == 0:
            e2e_render.readback(host_frames[0])  # the combined image
    for fence in fences:
        fence.wait()
    ctx.synchronize()
    dt_e2e = time.perf_counter() - t0
    e2e_ms, e2e_rays = gather_max_sum(1e3 * dt_e2e, int(ctx.stats().rays))
    e2e = {"value": e2e_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 400 + 48,
           "d2h_bytes_per_step": WIDTH * HEIGHT * 4, "ms_per_step": e2e_ms / n_e2e, "steps": n_e2e,
           "frames_in_flight": 2,
           "includes": "uniform block + params H2D, rgba8 frame D2H into page-locked memory every step (2 frames in flight, per-slot fences)" + ("; closing NCCL reduce + resolve + D2H of the combined image on rank 0" if world > 1 else "")}
    del e2e_accum, e2e_render, host_frames, fences

    roofline = None
    cpu_baseline = None
    extra = {}

    # ---- N > 1: the hard case of SURVEY 8e beside the frames split: ONE frame cut across the GPUs (interleaved 8-row bands,
    #      one all-gather per frame), against the same frame traced by one GPU alone in the same run ----
    if world > 1 and not tiled and not args.no_extra:
        n_t = 32  # short frames: enough of them that the fill and drain of the frame pipeline are a small part of the clock
        alone = device_timed(sd, cam, sbt, sched, 2, n_t, None, 300)       # every rank traces whole frames on its own
        tiles = device_timed(sd, cam, sbt, sched, 2, n_t, "tiles", 300)
        t_ms, _ = gather_max_sum(tiles["ms_total"], 0)
        a_ms, _ = gather_max_sum(alone["ms_total"], 0)
        extra["tile_split"] = {"ms_per_frame": t_ms / n_t, "single_gpu_ms_per_frame": a_ms / n_t, "speedup_vs_1gpu": a_ms / t_ms,
                               "frames": n_t, "band_rows": BAND,
                               "what": "one frame cut into interleaved %d-row bands over %d GPUs, solb_allgather_rows (peer stores into the IPC-mapped staging of every rank over NVLink, flags, scatter; ncclAllGather when the peers cannot be mapped) of the rgba8 frame per frame and of the float accumulation once at the end, all inside the timed region; max over ranks" % (BAND, world)}
        # configs[3] (3840 x 2160) with the frames split at this N: frames f = rank (mod N), one reduce + resolve at the end
        n_4k = 4
        r4 = device_timed(sd, cam, sbt, sched, 2, n_4k, "frames", 400, w=3840, h=2160)
        k_ms, k_rays = gather_max_sum(r4["ms_total"], r4["rays"])
        extra["config3_tunnel_4k"] = {"Mrays_s": k_rays / (k_ms * 1e-3) / 1e6, "ms_per_frame_per_gpu": (k_ms - r4["reduce_ms"]) / n_4k,
                                      "reduce_ms": r4["reduce_ms"], "frames_per_rank": n_4k, "resolution": "3840x2160", "n_gpus": world}
        cam.set_window_size((WIDTH, HEIGHT))

    if rank == 0:
        wkey = "tunnel_%dx%d" % (WIDTH, HEIGHT) if args.workload == "tunnel" else "synth_%d" % args.blas
        roofline = measure_roofline(sd, cam, sbt, sched, WIDTH, HEIGHT, MAX_BOUNCES, wkey)

    if rank == 0 and world == 1 and args.workload == "tunnel" and not args.size and not args.no_extra:
        # ---- the other BASELINE.json configs, each with its own roofline (bytes per ray from that run's counters) ----
        def line(r, n, sd_, w, h):
            return {"Mrays_s": r["rays"] / (r["ms_total"] * 1e-3) / 1e6, "ms_per_frame": r["ms_total"] / n,
                    "rays_per_path": r["rays"] / max(r["paths"], 1), "resolution": "%dx%d" % (w, h),
                    "schedule": sched_names[resolved_schedule(sd_, N.SCHEDULE_AUTO)]}

        # cornell beside tunnel (the metric names both): same resolution / spp / bounce cap
        _, sd_c, cam_c, sbt_c = setup("cornell.gltf", False)
        n_c = min(steps, 8)
        rc = device_timed(sd_c, cam_c, sbt_c, N.SCHEDULE_AUTO, 3, n_c, None, 0)
        extra["cornell_1080p_cap8"] = line(rc, n_c, sd_c, WIDTH, HEIGHT)
        extra["cornell_1080p_cap8"]["roofline"] = measure_roofline(sd_c, cam_c, sbt_c, N.SCHEDULE_AUTO, WIDTH, HEIGHT, MAX_BOUNCES, "cornell_1920x1080")
        # configs[3] on one GPU: 3840 x 2160
        r4 = device_timed(sd, cam, sbt, sched, 2, 4, None, 0, w=3840, h=2160)
        extra["config3_tunnel_4k"] = line(r4, 4, sd, 3840, 2160)
        extra["config3_tunnel_4k"]["roofline"] = measure_roofline(sd, cam, sbt, sched, 3840, 2160, MAX_BOUNCES, "tunnel_3840x2160")
        cam.set_window_size((WIDTH, HEIGHT))
        # configs[1]: 4-ray-ao, Duck.gltf stands in for ToyCar.glb (missing upstream), the example's camera, 1 primary + <= 4 AO rays x 4 samples
        sc_d = scene.load_scene(ctx, os.path.join(ROOT, "assets", "models", "Duck.gltf"))
        sd_d = ray.SceneDescription.from_scene(ctx, sc_d)
        cam_d = scene.Camera((WIDTH, HEIGHT))
        cam_d.look_at((4, 1, 4), (0, 0.5, 0), (0, -1, 0))  # examples/4-ray-ao.rs:89-90
        ctx.set_blue_noise(sol_io.load_texture_rgba8(os.path.join(ROOT, "assets", "textures", "HDR_RGBA_0.png")))
        pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/ao.rgen", ray.RAYGEN_KHR).shader("glsl/ao.rmiss", ray.MISS_KHR)
                            .shader("glsl/ao.rchit", ray.CLOSEST_HIT_KHR))
        sbt_ao = ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))
        ra = device_timed(sd_d, cam_d, None, None, 3, 8, None, 0,
                          trace_fn=lambda f, img: sbt_ao.cmd_trace_rays(ray.TraceBindings(sd_d, scene.scene_uniforms(cam_d, WIDTH, HEIGHT, f), img, None), (WIDTH, HEIGHT, 1)))
        extra["config1_ao_duck_1080p"] = {"Mrays_s": ra["rays"] / (ra["ms_total"] * 1e-3) / 1e6, "ms_per_frame": ra["ms_total"] / 8,
                                          "rays_per_frame": ra["rays"] / 8, "kernel": "k_ao",
                                          "note": "Duck.gltf stands in for the example's ToyCar.glb (absent upstream: .MISSING_LARGE_BLOBS)"}
        del sd_d, sc_d
        # configs[4]: synthetic 1000 BLAS x 20 000 triangles = 20 M triangles: BVH build + path trace on this GPU
        try:
            from sol_rs_b200 import synth

            t0 = time.perf_counter()
            sc_s = synth.make_scene(1000, 100)
            gen_s = time.perf_counter() - t0
            sd_s = ray.SceneDescription.from_scene(ctx, sc_s)
            first_build_ms = ctx.stats().last_build_ms
            builds = []
            for _ in range(3):  # rebuilds with the scratch pool warm (the figure BASELINE.md quotes)
                sd_s.blas_transform(sc_s.meshes[0].transform, 0)
                sd_s.tlas_regenerate()
                builds.append(ctx.stats().last_build_ms)
            info = sd_s.accel_info()
            sbt_s = pt_sbt(True)
            rs = device_timed(sd_s, sc_s.camera, sbt_s, sched, 2, 4, None, 0)
            extra["config4_synth_20m"] = line(rs, 4, sd_s, WIDTH, HEIGHT)
            extra["config4_synth_20m"].update({
                "triangles": int(info.n_triangles), "wide_nodes": int(info.n_wide_nodes), "wide_depth": int(info.wide_depth),
                "bvh_build_ms": min(builds), "bvh_build_ms_first": first_build_ms, "Mtris_s": info.n_triangles / (min(builds) * 1e-3) / 1e6,
                "build_roofline": {"bound": "hbm", "bytes_per_triangle": 379, "achieved": 379 * info.n_triangles / (min(builds) * 1e-3) / 1e9,
                                   "peak": measured_peak_gbs()[0], "unit": "GB/s",
                                   "frac": 379 * info.n_triangles / (min(builds) * 1e-3) / 1e9 / measured_peak_gbs()[0]},
                "sah": float(info.sah_cost_binary), "scene_gen_s": gen_s,
                "roofline": measure_roofline(sd_s, sc_s.camera, sbt_s, sched, WIDTH, HEIGHT, MAX_BOUNCES, "synth_1000")})
            del sd_s, sc_s
            ctx.trim()
        except Exception as e:  # e.g. a smaller GPU: the headline line must still be printed
            extra["config4_synth_20m"] = {"error": repr(e)}
    if rank == 0 and world == 1 and args.extra:
        for name, s_ in (("wavefront", N.SCHEDULE_WAVEFRONT), ("megakernel", N.SCHEDULE_MEGAKERNEL), ("warpfront", N.SCHEDULE_WARPFRONT)):
            r2 = device_timed(sd, cam, sbt, s_, 2, min(steps, 6), None, 0)
            extra["tunnel_" + name] = {"Mrays_s": r2["rays"] / (r2["ms_total"] * 1e-3) / 1e6, "ms_per_step": r2["ms_total"] / min(steps, 6)}
        _, sd_c, cam_c, sbt_c = setup("cornell.gltf", False)
        for name, s_ in (("wavefront", N.SCHEDULE_WAVEFRONT), ("megakernel", N.SCHEDULE_MEGAKERNEL), ("warpfront", N.SCHEDULE_WARPFRONT)):
            rc = device_timed(sd_c, cam_c, sbt_c, s_, 2, min(steps, 6), None, 0)
            extra["cornell_1080p_cap8_" + name] = {"Mrays_s": rc["rays"] / (rc["ms_total"] * 1e-3) / 1e6, "ms_per_step": rc["ms_total"] / min(steps, 6)}
    if rank == 0 and not args.no_cpu_baseline and world == 1 and args.workload == "tunnel" and not args.size:
        cb = cpu_reference_arm(1, 0, (WIDTH, HEIGHT))  # ONE full frame of the workload (~20 s on 16 host threads)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0 and roofline:
        # the same algorithmic bytes over the steady-state step of the timed region (frames in flight overlap one frame's drain
        # with the next frame's start; `achieved` / `frac` above are from launches timed one at a time)
        roofline["achieved_in_timed_region"] = roofline["bytes_per_ray"] * (rays_total / world) / (ms_total * 1e-3) / 1e9
        roofline["frac_in_timed_region"] = roofline["achieved_in_timed_region"] / roofline["peak"]
    if rank == 0:
        line_ = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                 "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if tiled else "weak", "vs_baseline": None,
                 "dtype": "f32", "data": "synthetic",
                 "config": {"workload": workload, "schedule": sched_names[resolved_schedule(sd, sched)], "accel": args.accel, "bvh_build_ms": build_ms, "bvh_rebuild_ms": min(rebuilds),
                            "frames_per_rank": steps, "rays_per_frame": rays_total / (steps * (1 if tiled else world)),
                            "l2": "256 MB buffer written between timed steps; the clock is one CUDA-event interval over all K steps, flushes included",
                            "frames_in_flight": int(os.environ.get("SOLB_WL_FRAMES_IN_FLIGHT", 3)) if sched_names[resolved_schedule(sd, sched)] == "warpfront" else 1,
                            "tlas": "solb_tlas_regenerate is called every frame like the reference (examples/5-pathtrace.rs:316) and is a no-op while no transform changed; "
                                    "a forced full rebuild of this scene costs bvh_build_ms",
                            "multi_gpu": ("every frame cut into 8-row bands dealt round-robin to the N ranks, one solb_allgather_rows (NCCL) of the rgba8 frame per frame + one of the accumulation at the end, inside the timed region"
                                          if tiled else
                                          "frames f = rank (mod N) per rank, local sums, one solb_reduce_accum (NCCL reduce + resolve on rank 0) inside the timed region")
                            if world > 1 else "single GPU"},
                 "ms_per_frame": ms_per_step, "reduce_ms": main_run["reduce_ms"],
                 "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": main_run["launches"],
                 "clocks": clock_info, "extra": extra}
        _emit(line_)
    if dist:
        dist.barrier()
        if comm:
            comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
