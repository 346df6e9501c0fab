"""CPU tier: the library's per-thread device logic (builder, 8-wide traversal, watertight test, shading),
compiled for the host by tests/emu and stepped sequentially, against the oracle.  The same checks run on the
real kernels in test_gpu_parity.py; this tier catches logic errors without GPU time."""
import numpy as np
import pytest

import oracle
from oracle import camera as ocam

import emu_lib
from helpers import oracle_camera, oracle_scene


@pytest.mark.parametrize("name,w,h", [("cornell", 192, 192), ("tunnel", 320, 180), ("Duck", 300, 200)])
@pytest.mark.parametrize("passes", [0, 2])
def test_emu_primary_hits(name, w, h, passes):
    fs, osc = oracle_scene(name)
    es = emu_lib.EmuScene(fs, passes)
    u = ocam.scene_uniforms(oracle_camera(fs, name, w, h), w, h, 0)
    _, ids, _, flags = osc.debug(u, w, h)
    _, eids = es.debug(u, w, h)
    mism = np.any(ids != eids, axis=2)
    assert (mism & (flags == 0)).sum() == 0
    info = es.info()
    assert info["sah"] <= info["sah_lbvh"] * 1.0001 and info["depth"] <= 62


def test_emu_treelet_pass_improves_sah():
    fs, _ = oracle_scene("tunnel")
    a, b = emu_lib.EmuScene(fs, 0).info(), emu_lib.EmuScene(fs, 2).info()
    assert b["sah"] < 0.8 * a["sah"]


def test_emu_random_rays_and_stack_depth():
    fs, osc = oracle_scene("Duck")
    es = emu_lib.EmuScene(fs, 2)
    rng = np.random.default_rng(5)
    n = 100_000
    lo, hi = osc.bounds()
    c, r = (lo + hi) / 2, np.linalg.norm(hi - lo)
    o = c + rng.normal(size=(n, 3)) * r
    d = (c + rng.normal(size=(n, 3)) * 0.2 * r) - o
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)
    hits, t, ctr = es.trace_rays(rays)
    o_hits, o_t, flags = osc.trace_rays(rays, classify=True)
    mism = np.any(hits[:, :2] != o_hits[:, :2], axis=1)
    assert (mism & (flags == 0)).sum() == 0
    assert (o_hits[:, 0] != oracle.MISS).mean() > 0.05
    assert es.info()["max_stack"] <= es.info()["depth"]


@pytest.mark.parametrize("name,w,h,sky,mb", [("cornell", 64, 64, False, 32), ("tunnel", 96, 54, True, 8)])
def test_emu_pathtrace_frame(name, w, h, sky, mb):
    fs, osc = oracle_scene(name)
    es = emu_lib.EmuScene(fs, 2)
    cam = oracle_camera(fs, name, w, h)
    a = np.zeros((h, w, 4), np.float32)
    b = np.zeros((h, w, 4), np.float32)
    st = oracle.OrcStats()
    for f in range(2):
        u = ocam.scene_uniforms(cam, w, h, f)
        osc.pathtrace_frame(u, w, h, a, 0, sky, 8, mb, st)
        es.pathtrace_frame(u, w, h, b, 0, sky, 8, mb)
    d = np.abs(a - b)[..., :3]
    assert (d.max(axis=2) > 1e-3 * (1 + a[..., :3].max(axis=2))).mean() < 0.02
    assert d.sum() / a[..., :3].sum() < 0.01


def test_emu_textured_duck_frame():
    """shade.cuh's base-colour texture path (an extension shared with the oracle) on the CPU tier: the device
    code compiled for the host, Duck.gltf with DuckCM.png, against the oracle carrying the same extension."""
    fs, osc = oracle_scene("Duck")
    assert len(fs.textures) == 1 and fs.material_textures == [0]
    es = emu_lib.EmuScene(fs, 2)
    osc.set_textures(fs.textures, fs.material_textures)
    es.set_textures(fs.textures, fs.material_textures, fs)
    w, h = 96, 72
    cam = oracle_camera(fs, "Duck", w, h)
    a, b, plain = (np.zeros((h, w, 4), np.float32) for _ in range(3))
    for f in range(2):
        u = ocam.scene_uniforms(cam, w, h, f)
        osc.pathtrace_frame(u, w, h, a, 0, True, 8, 4, oracle.OrcStats())
        es.pathtrace_frame(u, w, h, b, 0, True, 8, 4)
    d = np.abs(a - b)[..., :3]
    assert (d.max(axis=2) > 1e-3 * (1 + a[..., :3].max(axis=2))).mean() < 0.02
    assert d.sum() / a[..., :3].sum() < 0.01
    es.set_textures([], fs.material_textures, fs)
    for f in range(2):
        es.pathtrace_frame(ocam.scene_uniforms(cam, w, h, f), w, h, plain, 0, True, 8, 4)
    assert np.abs(plain - b)[..., :3].max(axis=2).mean() > 1e-3  # the texture did something


def test_emu_rng_bit_exact():
    L = emu_lib.lib()
    import ctypes
    rng = np.random.default_rng(0)
    for a, b in rng.integers(0, 2 ** 32, size=(500, 2)):
        assert L.emu_tea(int(a), int(b)) == oracle.tea(int(a), int(b))
    for seed in rng.integers(0, 2 ** 32, size=50):
        s = ctypes.c_uint32(int(seed))
        ref = oracle.rand_stream(int(seed), 8)
        for w, f in ref:
            assert L.emu_next_rand(ctypes.byref(s)) == f


# ---- two-level mode (TLAS over shared object-space BLASes) ------------------------------

@pytest.mark.parametrize("name,w,h", [("cornell", 192, 192), ("tunnel", 320, 180), ("Duck", 300, 200)])
def test_emu_two_level_primary_hits(name, w, h):
    fs, osc = oracle_scene(name)
    es = emu_lib.EmuScene(fs, 2, two_level=True)
    u = ocam.scene_uniforms(oracle_camera(fs, name, w, h), w, h, 0)
    _, ids, _, flags = osc.debug(u, w, h)
    _, eids = es.debug(u, w, h)
    mism = np.any(ids != eids, axis=2)
    assert (mism & (flags == 0)).sum() == 0
    assert (ids[..., 0] != oracle.MISS).mean() > 0.03


def _duck_rays(osc, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = osc.bounds()
    c, r = (lo + hi) / 2, np.linalg.norm(hi - lo)
    o = c + rng.normal(size=(n, 3)) * r
    d = (c + rng.normal(size=(n, 3)) * 0.2 * r) - o
    return np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)


def test_emu_two_level_shared_blas_instances():
    """four instances of ONE BLAS (mirrored, scaled, overlapping): two-level hits = oracle hits = flattened hits."""
    from helpers import duck_extras, instanced_variant

    fs, _ = oracle_scene("Duck")
    extras = duck_extras(fs)
    osc = oracle.Scene(instanced_variant(fs, extras))
    two = emu_lib.EmuScene(fs, 2, two_level=True, extra_instances=extras)
    flat = emu_lib.EmuScene(fs, 2, two_level=False, extra_instances=extras)
    rays = _duck_rays(osc, 60_000, 7)
    o_hits, o_t, flags = osc.trace_rays(rays, classify=True)
    for es in (two, flat):
        hits, t, _ = es.trace_rays(rays)
        mism = np.any(hits[:, :2] != o_hits[:, :2], axis=1)
        assert (mism & (flags == 0)).sum() == 0
        ok = (o_hits[:, 0] != oracle.MISS) & ~mism
        assert np.allclose(t[ok], o_t[ok], rtol=2e-4, atol=2e-5)
    assert set(np.unique(o_hits[:, 0])) >= {0, 1, 2, 3}  # every instance is hit
    # shared geometry: the two-level structure stores the duck once, the flattened one four times
    assert two.tris(4212).shape[0] == 4212 and flat.info()["nodes"] > 3 * (two.info()["nodes"] - 4)
    assert two.info()["max_stack"] <= 2 * two.info()["depth"] + 1


def test_emu_two_level_tlas_rebuild_after_transform_update():
    from helpers import duck_extras, instanced_variant, trs

    fs, _ = oracle_scene("Duck")
    extras = duck_extras(fs)
    two = emu_lib.EmuScene(fs, 2, two_level=True, extra_instances=extras)
    moved = np.ascontiguousarray((trs((0.5, 1.0, -2.0), (0, 1, 0), 2.0).T.astype(np.float64)
                                  @ np.asarray(extras[1][1], dtype=np.float64).T).T, dtype=np.float32)
    two.set_transform(2, moved)  # instance 2 = extras[1]
    two.build()
    extras2 = [extras[0], (0, moved, 0), extras[2]]
    osc = oracle.Scene(instanced_variant(fs, extras2))
    rays = _duck_rays(osc, 30_000, 11)
    o_hits, _, flags = osc.trace_rays(rays, classify=True)
    hits, _, _ = two.trace_rays(rays)
    assert (np.any(hits[:, :2] != o_hits[:, :2], axis=1) & (flags == 0)).sum() == 0
    assert (o_hits[:, 0] == 2).sum() > 100


@pytest.mark.parametrize("name,w,h,sky,mb", [("cornell", 64, 64, False, 32), ("tunnel", 96, 54, True, 8)])
def test_emu_two_level_pathtrace_frame(name, w, h, sky, mb):
    fs, osc = oracle_scene(name)
    es = emu_lib.EmuScene(fs, 2, two_level=True)
    cam = oracle_camera(fs, name, w, h)
    a = np.zeros((h, w, 4), np.float32)
    b = np.zeros((h, w, 4), np.float32)
    st = oracle.OrcStats()
    u = ocam.scene_uniforms(cam, w, h, 0)
    osc.pathtrace_frame(u, w, h, a, 0, sky, 8, mb, st)
    es.pathtrace_frame(u, w, h, b, 0, sky, 8, mb)
    d = np.abs(a - b)[..., :3]
    assert (d.max(axis=2) > 1e-3 * (1 + a[..., :3].max(axis=2))).mean() < 0.02
    assert d.sum() / a[..., :3].sum() < 0.01


def test_emu_optimal_collapse_beats_greedy():
    """SAH-optimal wide collapse (Ylitie et al. 2017) vs greedy largest-area-first on tunnel.gltf path-tracing rays:
    fewer wide nodes, no more node visits or triangle tests per ray, identical hits."""
    fs, _ = oracle_scene("tunnel")
    w, h = 160, 90
    u = ocam.scene_uniforms(oracle_camera(fs, "tunnel", w, h), w, h, 0)
    res = {}
    for dp in (True, False):
        es = emu_lib.EmuScene(fs, 2, dp_collapse=dp)
        acc = np.zeros((h, w, 4), np.float32)
        _, st = es.pathtrace_frame(u, w, h, acc, 0, True, 8, 8)
        res[dp] = (es.info()["nodes"], st[2] / st[0], st[3] / st[0], acc, es.debug(u, w, h)[1])
    assert res[True][0] < 0.8 * res[False][0]
    cost = lambda r: 2.0 * r[1] + 1.0 * r[2]
    assert cost(res[True]) < 0.99 * cost(res[False])
    assert np.array_equal(res[True][4], res[False][4])  # primary hit ids do not depend on the tree
    d = np.abs(res[True][3] - res[False][3])[..., :3]
    assert (d.max(axis=2) > 1e-3).mean() < 0.01


def test_node_hit_mask_table():
    """bvh.cuh: the per-node triangle-mask bytes, count0 and the per-ray octant bytes reproduce, through four byte dot products, the
    hit mask built child by child (internal child of slot s -> bit 24 + (s ^ oct_inv), leaf -> its run of triangle bits) for random
    child sets, random slab results and all eight octants; decode_child_kind returns what encode_node8 was given."""
    assert emu_lib.lib().emu_mask_selftest(0xB200, 20000) == 0
