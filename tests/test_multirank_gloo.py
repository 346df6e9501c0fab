"""CPU tier, world_size 2 over gloo: the multi-GPU split of SURVEY 8e (frames f = r mod R per rank, per-rank
sums, one sum-reduce, resolve on rank 0) reproduces the single-process running mean of pathtrace.rgen:89-101.
Per-frame colours come from the oracle here (no GPU); the GPU path uses the same host logic with nccl."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, N_FRAMES = 48, 32, 6


def _frame_colour(osc, cam, f):
    from oracle import camera as ocam

    img = np.zeros((H, W, 4), dtype=np.float32)
    osc.pathtrace_frame(ocam.scene_uniforms(cam, W, H, f), W, H, img, f, False, 8, 4)  # start = f -> alpha = 1: this frame only
    return img


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import oracle_camera, oracle_scene
    from sol_rs_b200 import multigpu

    fs, osc = oracle_scene("cornell")
    cam = oracle_camera(fs, "cornell", W, H)
    acc = torch.zeros((H, W, 4), dtype=torch.float32)
    mine = multigpu.frames_for_rank(rank, world, N_FRAMES)
    for f in mine:
        c = _frame_colour(osc, cam, f)
        acc[..., :3] += torch.from_numpy(c[..., :3])
        acc[..., 3] += 1.0
    multigpu.reduce_accum(acc, dst=0)
    # tile split + allgather (single-sample interactive mode): interleaved bands, ragged last band (H = 32, bands of 5 rows)
    for band in (5, 8):
        img = torch.full((H, W, 4), -1.0)
        for y in multigpu.bands_for_rank(rank, world, H, band):
            img[y] = float(1000 * rank + y)
        full = multigpu.allgather_band_rows(img, rank, world, band)
        want = torch.tensor([float(1000 * ((y // band) % world) + y) for y in range(H)])
        assert full.shape == (H, W, 4) and torch.equal(full[:, 0, 0], want) and torch.equal(full[:, -1, 3], want)
    if rank == 0:
        np.save(out_path, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_frame_partition():
    from sol_rs_b200 import multigpu

    for world in (1, 2, 4, 8):
        owned = [multigpu.frames_for_rank(r, world, 512) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(512))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) == 0
    assert multigpu.frames_for_rank(1, 2, 5, first=10) == [11, 13]
    with pytest.raises(ValueError):
        multigpu.frames_for_rank(2, 2, 4)
    # tile split: the ranks' bands partition the image rows; the C ABI's (begin, count, stride) triple names the same rows
    for world in (1, 2, 3, 8):
        for height, band in ((1080, 8), (33, 5), (7, 8)):
            rows = [multigpu.bands_for_rank(r, world, height, band) for r in range(world)]
            assert sorted(sum(rows, [])) == list(range(height))
            for r in range(world):
                b0, cnt, stride = multigpu.tile_rows_for_rank(r, world, band)
                named = [y for k in range(0, height, stride) for y in range(b0 + k, min(b0 + k + cnt, height))]
                assert named == rows[r]


def test_two_rank_sum_reduce_equals_running_mean(tmp_path):
    out = str(tmp_path / "acc.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    summed = np.load(out)
    assert np.all(summed[..., 3] == N_FRAMES)
    # single-process reference: the reference's running mix over frames 0..N-1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_camera, oracle_scene
    from oracle import camera as ocam

    fs, osc = oracle_scene("cornell")
    cam = oracle_camera(fs, "cornell", W, H)
    mix = np.zeros((H, W, 4), dtype=np.float32)
    for f in range(N_FRAMES):
        osc.pathtrace_frame(ocam.scene_uniforms(cam, W, H, f), W, H, mix, 0, False, 8, 4)
    np.testing.assert_allclose(summed[..., :3] / N_FRAMES, mix[..., :3], rtol=2e-5, atol=1e-6)
