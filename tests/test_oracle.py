"""Pins for the CPU oracle itself (no GPU).  The reference holds no fixtures, so the
pins are: exact-integer RNG KATs derived from sampling.glsl, scene-flatten counts /
bounds (8c (2)), camera fixtures (8c (3)) and the structural rays-per-path statistics."""
import ctypes
import os

import numpy as np
import pytest

import oracle
from oracle import camera as ocam
from oracle import gltf_flatten as gf


def test_tea_kats():
    # assets/glsl/sampling.glsl:18-32
    assert oracle.tea(0, 0) == 0x741C187D
    assert oracle.tea(1, 0) == 0x8DA6B311
    assert oracle.tea(0, 1) == 0x70D3AEF1
    assert oracle.tea(2073599, 0) == 0xCDA8C132
    assert oracle.tea(2073599, 511) == 0x64160DDE
    assert oracle.tea(130815, 7) == 0x65D41AC3


def test_next_rand_kats():
    # assets/glsl/sampling.glsl:35-42
    s = oracle.rand_stream(oracle.tea(0, 0), 3)
    assert [w for w, _ in s] == [0x9999C9A7, 0x87F0A1F8, 0xA688239B]
    np.testing.assert_allclose([f for _, f in s], [0.6000028849, 0.5310155153, 0.6505148411], rtol=0, atol=1e-9)
    for w, f in s:
        assert f == np.float32(w) * np.float32(2.0 ** -32)


def test_next_rand_reaches_one():
    # float(0xffffff80) / 4294967295.0f == 1.0f exactly (SURVEY App.A 11): find a state producing a big word
    assert np.float32(0xFFFFFF80) / np.float32(4294967295.0) == np.float32(1.0)
    assert np.float32(0xFFFFFF7F) / np.float32(4294967295.0) < np.float32(1.0)


def test_tea_matches_python_twin():
    def tea(v0, v1):
        s0 = 0
        M = 0xFFFFFFFF
        for _ in range(16):
            s0 = (s0 + 0x9E3779B9) & M
            v0 = (v0 + ((((v1 << 4) & M) + 0xA341316C) ^ ((v1 + s0) & M) ^ ((v1 >> 5) + 0xC8013EA4))) & M
            v1 = (v1 + ((((v0 << 4) & M) + 0xAD90777D) ^ ((v0 + s0) & M) ^ ((v0 >> 5) + 0x7E95761E))) & M
        return v0
    rng = np.random.default_rng(1)
    for a, b in rng.integers(0, 2 ** 32, size=(200, 2)):
        assert oracle.tea(int(a), int(b)) == tea(int(a), int(b))


@pytest.mark.parametrize("name,n_inst,n_tris,n_verts,lo,hi", [
    ("cornell", 8, 32, 64, (-1.02, 0.0, -1.04), (1.0, 1.99, 0.99)),
    ("tunnel", 2, 13116, 39348, (-1.0, -1.0, -7.956), (1.0, 1.0, 5.2)),
    ("Duck", 1, 4212, 2399, (-0.693, 0.099, -0.613), (0.962, 1.64, 0.539)),
])
def test_flatten_fixtures(assets, name, n_inst, n_tris, n_verts, lo, hi):
    fs = gf.load_scene(os.path.join(assets, "models", name + ".gltf"))
    sc = oracle.Scene(fs)
    assert len(fs.instances) == n_inst
    assert sc.tri_count == n_tris
    assert fs.vertices.shape == (n_verts, 16)
    blo, bhi = sc.bounds()
    np.testing.assert_allclose(blo, lo, atol=2e-3)
    np.testing.assert_allclose(bhi, hi, atol=2e-3)
    # instance ids are the running (mesh x primitive) count; indices are section-relative
    for inst in fs.instances:
        idx = fs.indices[inst["first_index"]: inst["first_index"] + inst["n_indices"]]
        assert idx.max() < inst["n_vertices"]


def test_flatten_details(assets):
    t = gf.load_scene(os.path.join(assets, "models", "tunnel.gltf"))
    assert [i["n_indices"] // 3 for i in t.instances] == [9120, 3996]
    assert [i["material"] for i in t.instances] == [0, 1]
    for inst in t.instances:  # mirrored X (det -1), SURVEY hard part 5
        assert np.linalg.det(inst["transform"][:3, :3].astype(np.float64)) == pytest.approx(-1.0)
    # gltf defaults: base colour 1, metallic given 0, roughness .3/.85, emissive 0
    np.testing.assert_allclose(t.materials[0], [1, 1, 1, 1, 0, 0, 0, 0, 0, 0.3, 0, 0], atol=1e-7)
    np.testing.assert_allclose(t.materials[1][9], 0.85, atol=1e-7)
    # COLOR_0 is vec3 -> alpha 1; normal.w = 1; pos.w = 1
    assert np.all(t.vertices[:, 3] == 1) and np.all(t.vertices[:, 7] == 1) and np.all(t.vertices[:, 11] == 1)
    c = gf.load_scene(os.path.join(assets, "models", "cornell.gltf"))
    assert [i["n_indices"] // 3 for i in c.instances] == [2, 2, 2, 2, 2, 2, 10, 10]
    np.testing.assert_allclose(c.materials[0][4:7], [15, 15, 15])
    assert c.materials[7][8] == np.float32(0.9999999776482582) and c.materials[7][9] == 0
    d = gf.load_scene(os.path.join(assets, "models", "Duck.gltf"))
    # hierarchy depth 2: root scale 0.01 applied to the child mesh node
    np.testing.assert_allclose(np.diag(d.instances[0]["transform"])[:3], 0.01, rtol=1e-6)
    assert np.all(d.vertices[:, 4:8] == 1)  # no COLOR_0 -> white


def test_camera_fixtures(assets):
    c = gf.load_scene(os.path.join(assets, "models", "cornell.gltf"))
    cam = ocam.Camera.from_view(c.camera["view"], c.camera["yfov"], c.camera["znear"], c.camera["zfar"])
    cam.set_window_size((512, 512))
    u = np.frombuffer(ocam.scene_uniforms(cam, 512, 512, 3), dtype=np.float32, count=96).reshape(6, 4, 4)
    vi = u[2]
    np.testing.assert_allclose(vi[3, :3], [0, 1, 4.1], atol=1e-6)       # origin
    np.testing.assert_allclose(-vi[2, :3], [0, 0, -1], atol=1e-6)      # forward
    np.testing.assert_allclose(vi[1, :3], [0, -1, 0], atol=1e-6)       # view-up -> world
    assert max(1.0, np.linalg.norm(vi[3, :3])) * 1e-3 == pytest.approx(4.22e-3, rel=1e-3)
    # closed-form proj_inv: xyz of proj_inv*(dx,dy,1,1) is proportional to (dx/w, dy/h, -1)
    pi = u[4]
    h = 1.0 / np.tan(np.radians(35.0) / 2)
    v = pi.T @ np.array([0.3, -0.7, 1, 1], dtype=np.float32)
    np.testing.assert_allclose(v[:3] / -v[2], [0.3 / h, -0.7 / h, -1], rtol=1e-5)
    frame = np.frombuffer(ocam.scene_uniforms(cam, 512, 512, 3), dtype=np.uint32, count=100)[96:99]
    assert list(frame) == [512, 512, 3]
    t = gf.load_scene(os.path.join(assets, "models", "tunnel.gltf"))
    cam = ocam.Camera.from_view(t.camera["view"], t.camera["yfov"], t.camera["znear"], t.camera["zfar"])
    cam.set_window_size((1920, 1080))
    vi = np.frombuffer(ocam.scene_uniforms(cam, 1920, 1080, 0), dtype=np.float32, count=96).reshape(6, 4, 4)[2]
    np.testing.assert_allclose(vi[3, :3], [0, -0.8, 0], atol=1e-6)
    np.testing.assert_allclose(-vi[2, :3], [0, 0.158, -0.9874], atol=1e-4)


def test_mat4_inverse_against_f64():
    rng = np.random.default_rng(7)
    for _ in range(20):
        m = rng.normal(size=(4, 4)).astype(np.float32)
        inv = gf.mat4_inverse(m)
        np.testing.assert_allclose(inv.astype(np.float64), np.linalg.inv(m.astype(np.float64)), rtol=2e-3, atol=2e-4)


def test_look_at_debug_camera():
    # examples/3-ray-debug.rs:78-79
    cam = ocam.Camera((900, 600))
    cam.look_at((5, 5, 5), (0, 0, 0), (0, -1, 0))
    vi = gf.mat4_inverse(cam.view)
    np.testing.assert_allclose(vi[3, :3], [5, 5, 5], atol=1e-5)
    np.testing.assert_allclose(-vi[2, :3], -np.ones(3) / np.sqrt(3), atol=1e-6)


def test_closest_hit_brute_force(assets):
    """The oracle's BVH must agree with a numpy brute force over all triangles."""
    fs = gf.load_scene(os.path.join(assets, "models", "cornell.gltf"))
    sc = oracle.Scene(fs)
    rng = np.random.default_rng(3)
    n = 2000
    o = rng.uniform([-1, 0, -1], [1, 2, 4], size=(n, 3))
    d = rng.normal(size=(n, 3))
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)
    hits, t, _ = sc.trace_rays(rays)
    # brute force in f64
    tris = []
    for ii, inst in enumerate(fs.instances):
        M = inst["transform"].astype(np.float64)
        for p in range(inst["n_indices"] // 3):
            idx = fs.indices[inst["first_index"] + 3 * p: inst["first_index"] + 3 * p + 3] + inst["first_vertex"]
            P = fs.vertices[idx, 0:3].astype(np.float64)
            tris.append((ii, p, P @ M[:3, :3] + M[3, :3]))  # [col][row] storage: world = sum_k M[k]*p[k] + M[3]
    best = np.full(n, np.inf)
    bid = np.full((n, 2), 0xFFFFFFFF, dtype=np.uint32)
    O = rays[:, 0:3].astype(np.float64)
    D = rays[:, 4:7].astype(np.float64)
    for ii, p, P in tris:
        e1, e2 = P[1] - P[0], P[2] - P[0]
        pv = np.cross(D, e2)
        det = pv @ e1
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            tv = O - P[0]
            uu = np.einsum("ij,ij->i", tv, pv) * inv
            q = np.cross(tv, e1)
            vv = np.einsum("ij,ij->i", D, q) * inv
            tt = (q @ e2) * inv
        ok = (uu >= 0) & (vv >= 0) & (uu + vv <= 1) & (tt > 1e-3) & (tt < 1e4) & (tt < best)
        best[ok] = tt[ok]
        bid[ok] = (ii, p)
    agree = np.all(hits[:, :2] == bid, axis=1)
    assert agree.mean() > 0.999  # exact ties / edge-on rays may differ
    hit = bid[:, 0] != 0xFFFFFFFF
    np.testing.assert_allclose(t[hit & agree], best[hit & agree], rtol=1e-6)


def test_rays_per_path_statistics(assets):
    """Structural check against SURVEY 6: tunnel --sky cap 32 ~ 15 rays/path, ~26 % capped."""
    fs = gf.load_scene(os.path.join(assets, "models", "tunnel.gltf"))
    sc = oracle.Scene(fs)
    cam = ocam.Camera.from_view(fs.camera["view"], fs.camera["yfov"], fs.camera["znear"], fs.camera["zfar"])
    W, H = 96, 54
    cam.set_window_size((W, H))
    acc = np.zeros((H, W, 4), dtype=np.float32)
    st = oracle.OrcStats()
    sc.pathtrace_frame(ocam.scene_uniforms(cam, W, H, 0), W, H, acc, 0, True, 8, 32, st)
    assert st.paths == W * H * 8
    assert 13.5 < st.rays / st.paths < 16.5
    assert 0.20 < st.capped / st.paths < 0.32
    assert np.all(np.isfinite(acc)) and np.all(acc[..., 3] == 1)


def test_accumulation_alpha_and_restart(assets):
    """pathtrace.rgen:89-101: frame f with start s uses alpha = 1/(f+1-s); alpha = 1 overwrites."""
    fs = gf.load_scene(os.path.join(assets, "models", "cornell.gltf"))
    sc = oracle.Scene(fs)
    cam = ocam.Camera.from_view(fs.camera["view"], fs.camera["yfov"], fs.camera["znear"], fs.camera["zfar"])
    W = H = 32
    cam.set_window_size((W, H))
    a = np.full((H, W, 4), 123.0, dtype=np.float32)
    sc.pathtrace_frame(ocam.scene_uniforms(cam, W, H, 5), W, H, a, 5, False, 8, 4)
    b = np.zeros((H, W, 4), dtype=np.float32)
    sc.pathtrace_frame(ocam.scene_uniforms(cam, W, H, 5), W, H, b, 5, False, 8, 4)
    np.testing.assert_array_equal(a, b)  # alpha == 1: old * 0 + new
    c = b.copy()
    sc.pathtrace_frame(ocam.scene_uniforms(cam, W, H, 6), W, H, c, 5, False, 8, 4)
    f6 = np.zeros((H, W, 4), dtype=np.float32)
    sc.pathtrace_frame(ocam.scene_uniforms(cam, W, H, 6), W, H, f6, 6, False, 8, 4)
    np.testing.assert_allclose(c[..., :3], b[..., :3] * np.float32(0.5) + f6[..., :3] * np.float32(0.5), rtol=1e-6, atol=1e-7)


# ---- committed golden fixtures (tests/golden/make_golden.py): the oracle must keep reproducing them ----------

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,cam_name,w,h", [("cornell", "cornell", 256, 256), ("Duck", "Duck", 450, 300), ("tunnel", "tunnel", 480, 270)])
def test_golden_hit_ids(name, cam_name, w, h):
    from helpers import oracle_camera, oracle_scene

    g = np.load(os.path.join(GOLDEN, "hitids_%s_%dx%d.npz" % (name, w, h)))
    fs, osc = oracle_scene(name)
    _, ids, _, flags = osc.debug(ocam.scene_uniforms(oracle_camera(fs, cam_name, w, h), w, h, 0), w, h)
    inst = np.where(ids[..., 0] == oracle.MISS, 255, ids[..., 0]).astype(np.uint8)
    prim = np.where(ids[..., 1] == oracle.MISS, 65535, ids[..., 1]).astype(np.uint16)
    assert np.array_equal(inst, g["inst"]) and np.array_equal(prim, g["prim"]) and np.array_equal(flags, g["flags"])


def test_golden_ao_frames():
    from helpers import load_blue_noise, oracle_camera, oracle_scene

    g = np.load(os.path.join(GOLDEN, "ao_duck_160x90_f0-3.npz"))
    fs, osc = oracle_scene("Duck")
    cam = oracle_camera(fs, "Duck_ao", 160, 90)
    img = np.zeros((90, 160, 4), dtype=np.float32)
    st = oracle.OrcStats()
    blue = load_blue_noise()
    assert blue.shape == (256, 256, 4) and blue.dtype == np.uint8
    for f in range(4):
        osc.ao_frame(ocam.scene_uniforms(cam, 160, 90, f), 160, 90, img, blue, 0, st)
    np.testing.assert_allclose(img[..., :3], g["image"], rtol=0, atol=1e-6)
    assert st.rays == int(g["rays"]) and st.paths == int(g["paths"])
    assert 0.0 <= img[..., :3].min() and img[..., :3].max() <= 1.0


def test_golden_converged_statistics():
    """The 4096-spp fixtures carry the structural statistics of SURVEY 6."""
    t = np.load(os.path.join(GOLDEN, "converged_tunnel_160x90_4096spp_b8.npz"))
    assert t["accum"].shape == (90, 160, 3) and int(t["paths"]) == 160 * 90 * 4096
    assert 5.9 < int(t["rays"]) / int(t["paths"]) < 6.7            # cap 8: ~6.3 rays/path
    assert 0.45 < int(t["capped"]) / int(t["paths"]) < 0.57         # ~51 % reach the cap
    c = np.load(os.path.join(GOLDEN, "converged_cornell_96x96_4096spp_b32.npz"))
    assert c["accum"].shape == (96, 96, 3) and np.all(np.isfinite(c["accum"]))
    assert 0.01 < int(c["emissive"]) / int(c["paths"]) < 0.06


@pytest.mark.parametrize("name,w,h,sky,mb,frame", [("cornell", 12, 12, False, 32, 3), ("tunnel", 16, 9, True, 8, 0), ("Duck", 12, 8, True, 8, 7)])
def test_oracle_matches_second_glsl_restatement(name, w, h, sky, mb, frame):
    """oracle.c against tests/glsl_twin.py, a second statement-by-statement reading of pathtrace.rgen / rchit / rmiss and
    sampling.glsl in numpy float32 (closest hits taken from the oracle's own trace_rays): same number of rays, i.e. the same
    RNG consumption and the same BRDF branch at every bounce, and the same pixel values to float rounding."""
    import glsl_twin
    from helpers import oracle_camera, oracle_scene

    fs, osc = oracle_scene(name)
    cam = oracle_camera(fs, name, w, h)
    u = ocam.scene_uniforms(cam, w, h, frame)
    twin, n_rays = glsl_twin.render_frame(fs, osc, u, w, h, sky, 8, mb)
    acc = np.zeros((h, w, 4), dtype=np.float32)
    st = oracle.OrcStats()
    osc.pathtrace_frame(u, w, h, acc, frame, sky, 8, mb, st)
    assert n_rays == st.rays
    d = np.abs(twin[..., :3] - acc[..., :3])
    assert d.max() <= 2e-4 * (1.0 + np.abs(acc[..., :3]).max())  # libm vs numpy sin / cos, amplified over the bounces
    assert np.all(acc[..., 3] == 1.0)


@pytest.mark.parametrize("name,cam_name,w,h,frame", [("Duck", "Duck_ao", 24, 14, 2), ("cornell", "cornell", 12, 12, 0)])
def test_oracle_ao_matches_second_glsl_restatement(name, cam_name, w, h, frame):
    """oracle.c's 4-ray-ao frame against tests/glsl_twin.py's reading of ao.rgen / ao.rchit / ao.rmiss (blue-noise lookup, chained
    cosine-weighted rays in [0.001, 10], hit counting from the second hit on): same rays, same pixels."""
    import glsl_twin
    from helpers import load_blue_noise, oracle_camera, oracle_scene

    fs, osc = oracle_scene(name)
    blue = load_blue_noise()
    u = ocam.scene_uniforms(oracle_camera(fs, cam_name, w, h), w, h, frame)
    twin, n_rays = glsl_twin.render_ao_frame(fs, osc, u, w, h, blue)
    img = np.zeros((h, w, 4), dtype=np.float32)
    st = oracle.OrcStats()
    osc.ao_frame(u, w, h, img, blue, frame, st)
    assert n_rays == st.rays
    assert np.abs(twin[..., :3] - img[..., :3]).max() <= 1e-6
    assert 0.0 < img[..., :3].mean() <= 1.0


# ---- fixtures from the real reference, when somebody has produced them (tools/reference_fixtures/README.md) ------------------
# The build image cannot run sol-rs (no rustc / Vulkan / lavapipe), so these files do not exist in the repository and the
# oracle stays "parity unpinned"; on a machine with the toolchain one command writes them and these tests pin the oracle.

def _ref_fixture(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("%s absent: run tools/reference_fixtures/run_reference_lavapipe.sh where sol-rs can run" % name)
    return np.load(path)


def test_reference_fixtures_primary_hit_ids():
    """3-ray-debug of the reference (id-writing stages) against the oracle's ids, outside the oracle's edge / tie list."""
    from helpers import oracle_camera, oracle_scene

    ref = _ref_fixture("ref_ids_Duck_900x600.npz")["ids"]
    fs, osc = oracle_scene("Duck")
    _, ids, _, flags = osc.debug(ocam.scene_uniforms(oracle_camera(fs, "Duck", 900, 600), 900, 600, 0), 900, 600)
    mism = np.any(ids != ref, axis=2)
    assert (mism & (flags == 0)).sum() == 0, "%d unlisted pixels differ from the reference" % (mism & (flags == 0)).sum()


@pytest.mark.parametrize("fixture,name,w,h", [("ref_frame_cornell_512x512_f8.npz", "cornell", 512, 512),
                                              ("ref_frame_tunnel_480x270_f8.npz", "tunnel", 480, 270)])
def test_reference_fixtures_accumulated_frames(fixture, name, w, h):
    """5-pathtrace of the reference after N accumulated frames (its own literals: 8 spp, 32 bounces) against the oracle's
    display image: same RNG streams, so PSNR > 40 dB unless the restatement is wrong."""
    from helpers import oracle_camera, oracle_scene

    ref = _ref_fixture(fixture)
    fs, osc = oracle_scene(name)
    cam = oracle_camera(fs, name, w, h)
    acc = np.zeros((h, w, 4), dtype=np.float32)
    rgba = None
    for f in range(int(ref["frames"])):
        rgba, _ = osc.pathtrace_frame(ocam.scene_uniforms(cam, w, h, f), w, h, acc, 0, bool(int(ref["sky"])), 8, 32)
    d = rgba[..., :3].astype(np.float64) - ref["rgba8"][..., :3].astype(np.float64)
    psnr = 10 * np.log10(255.0 ** 2 / max(np.mean(d ** 2), 1e-12))
    assert psnr > 40.0, "PSNR %.1f dB against the reference" % psnr
