"""GPU tier: the multi-GPU exchange behind the C ABI (csrc/comm.cu: solb_comm_*, solb_reduce_accum, solb_allgather_rows).
World 1 runs on any GPU box; the two-rank cases need two GPUs and are skipped otherwise (the N > 1 sharding logic itself
is covered on CPU by tests/test_multirank_gloo.py)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 256, 144


def _setup(sol, ctx):
    from helpers import model_path, pathtrace_pipeline
    from sol_rs_b200 import ray, scene

    sc = scene.load_scene(ctx, model_path("tunnel"))
    sd = ray.SceneDescription.from_scene(ctx, sc)
    cam = sc.camera
    cam.set_window_size((W, H))
    return sd, cam, pathtrace_pipeline(ctx, True)


def _frame(sol, ctx, sd, cam, sbt, f, accum, render=None, mode=0, tile=None):
    from sol_rs_b200 import ray, scene

    sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, W, H, f), accum, render, samples_per_frame=8, max_bounces=8,
                                         accum_mode=mode, tile_rows=tile), (W, H, 1))


def test_world_one_reduce_is_the_resolve_and_gather_is_a_noop():
    import sol_rs_b200 as sol
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import multigpu, ray

    ctx = sol.Context(0)
    sd, cam, sbt = _setup(sol, ctx)
    s = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F)
    for f in range(3):
        _frame(sol, ctx, sd, cam, sbt, f, s, mode=N.ACCUM_SUM)
    before = s.readback()
    a1, r1 = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F), sol.Image2d(ctx, W, H, N.FORMAT_RGBA8)
    a2, r2 = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F), sol.Image2d(ctx, W, H, N.FORMAT_RGBA8)
    ray.resolve_sum(ctx, s, a1, r1)
    h = lambda t: t.handle
    N.check(N.lib().solb_reduce_accum(ctx.handle, h(s), 0, h(a2), h(r2)), ctx.handle)  # no communicator: world 1
    assert np.array_equal(a1.readback(), a2.readback()) and np.array_equal(r1.readback(), r2.readback())
    assert np.all(before[..., 3] == 3.0) and np.array_equal(s.readback(), before)
    N.check(N.lib().solb_allgather_rows(ctx.handle, h(s), 8), ctx.handle)
    assert np.array_equal(s.readback(), before)
    with pytest.raises(sol.SolbError):
        N.check(N.lib().solb_reduce_accum(ctx.handle, h(s), 1, None, None), ctx.handle)  # root outside the world
    with pytest.raises(sol.SolbError):
        N.check(N.lib().solb_allgather_rows(ctx.handle, h(s), 0), ctx.handle)
    with pytest.raises(sol.SolbError):
        multigpu.Communicator(ctx, b"x" * 128, 2, 2)  # rank outside the world
    try:
        cid = multigpu.unique_id()
    except sol.SolbError:
        pytest.skip("no NCCL library in this process")
    assert len(cid) == 128
    ctx.close()


def _rank_main(rank, world, id_path, out_dir, p2p):
    os.environ["SOLB_P2P"] = "1" if p2p else "0"  # read once per process by libsolb
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import time

    import torch

    import sol_rs_b200 as sol
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import multigpu

    torch.cuda.set_device(rank)
    ctx = sol.Context(rank)
    # bootstrap through a file: the C ABI only needs the 128 bytes to reach every rank somehow
    if rank == 0:
        with open(id_path + ".tmp", "wb") as f:
            f.write(multigpu.unique_id())
        os.replace(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.01)
    comm = multigpu.Communicator.from_id(ctx, open(id_path, "rb").read(), rank, world)
    assert comm.info()["world"] == world and comm.info()["nccl_version"] > 20000
    sd, cam, sbt = _setup(sol, ctx)
    n_frames = 6
    # frames split: local sums, one reduce + resolve on rank 0
    s = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F)
    r = sol.Image2d(ctx, W, H, N.FORMAT_RGBA8)
    for f in multigpu.frames_for_rank(rank, world, n_frames):
        _frame(sol, ctx, sd, cam, sbt, f, s, mode=N.ACCUM_SUM)
    comm.reduce_accum(s, 0, s, r)
    # tile split: interleaved bands of 8 and (ragged against the 144 rows) 10 rows, all-gather after every frame
    tiles = {}
    for band in (8, 10):
        a = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F)
        for f in range(2):
            _frame(sol, ctx, sd, cam, sbt, f, a, tile=multigpu.tile_rows_for_rank(rank, world, band))
            comm.allgather_rows(a, band)
        tiles[band] = a.readback()
    # the way bench.py's tile split runs: many frames in flight, the rgba8 frame gathered after every frame (each read back
    # without waiting, two host slots), the float accumulation - every rank owns its bands of it - gathered once at the end
    a, rend = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F), sol.Image2d(ctx, W, H, N.FORMAT_RGBA8)
    slots, fences, shown = [ctx.host_alloc((H, W, 4), np.uint8) for _ in range(2)], [ctx.fence() for _ in range(2)], []
    for f in range(8):
        fences[f & 1].wait()
        if f >= 2:
            shown.append(slots[f & 1].copy())
        _frame(sol, ctx, sd, cam, sbt, f, a, rend, tile=multigpu.tile_rows_for_rank(rank, world, 8))
        comm.allgather_rows(rend, 8)
        rend.readback_async(slots[f & 1])
        fences[f & 1].signal()
    for f in (6, 7):
        fences[f & 1].wait()
        shown.append(slots[f & 1].copy())
    comm.allgather_rows(a, 8)
    extra = dict(t8=tiles[8], t10=tiles[10], shown=np.stack(shown), loop_accum=a.readback())
    if rank == 0:
        np.savez(os.path.join(out_dir, "rank0.npz"), accum=s.readback(), render=r.readback(), **extra)
    else:
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **extra)
    comm.close()
    ctx.close()


@pytest.mark.parametrize("p2p", [True, False], ids=["peer_stores", "nccl_allgather"])
def test_two_ranks_reduce_and_band_gather(tmp_path, p2p):
    """p2p: the band exchange as 64-thread CTAs storing into the peers' IPC-mapped staging blocks (default), or as pack +
    ncclAllGather + scatter (SOLB_P2P=0, and what a rank falls back to when a peer's block cannot be mapped)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    import sol_rs_b200 as sol
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray

    mp.spawn(_rank_main, args=(2, str(tmp_path / "nccl_id"), str(tmp_path), p2p), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # single-GPU references
    ctx = sol.Context(0)
    sd, cam, sbt = _setup(sol, ctx)
    s = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F)
    for f in range(6):
        _frame(sol, ctx, sd, cam, sbt, f, s, mode=N.ACCUM_SUM)
    a, r = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F), sol.Image2d(ctx, W, H, N.FORMAT_RGBA8)
    ray.resolve_sum(ctx, s, a, r)
    # the two-rank sum adds the same six frame colours in another order: equal to float rounding
    np.testing.assert_allclose(r0["accum"][..., :3], a.readback()[..., :3], rtol=1e-5, atol=1e-6)
    assert np.abs(r0["render"].astype(np.int32) - r.readback().astype(np.int32)).max() <= 1
    full = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F)
    for f in range(2):
        _frame(sol, ctx, sd, cam, sbt, f, full)
    want = full.readback()
    for key in ("t8", "t10"):  # the tile split is bit-identical to the undivided frames, on every rank
        assert np.array_equal(r0[key], want) and np.array_equal(r1[key], want), key
    full, rend = sol.Image2d(ctx, W, H, N.FORMAT_RGBA32F), sol.Image2d(ctx, W, H, N.FORMAT_RGBA8)
    for f in range(8):
        _frame(sol, ctx, sd, cam, sbt, f, full, rend)
        shown = rend.readback()
        assert np.array_equal(r0["shown"][f], shown) and np.array_equal(r1["shown"][f], shown), f
    want = full.readback()
    assert np.array_equal(r0["loop_accum"], want) and np.array_equal(r1["loop_accum"], want)
    ctx.close()
