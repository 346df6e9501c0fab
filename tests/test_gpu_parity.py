"""GPU parity tests: the CUDA path (through the C ABI, via sol_rs_b200) against the CPU oracle.
Run on the B200 box: python -m pytest tests -m gpu"""
import ctypes

import numpy as np
import pytest

import oracle
from oracle import camera as ocam

from helpers import (flat_from_product_scene, image_metrics, load_blue_noise, model_path, oracle_camera, oracle_scene,
                     pathtrace_pipeline, product_camera, simple_pipeline)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sol():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sol_rs_b200

    return sol_rs_b200


@pytest.fixture(scope="module")
def ctx(sol):
    c = sol.Context(0)
    yield c
    c.close()


def _product(sol, ctx, name):
    from sol_rs_b200 import ray, scene

    sc = scene.load_scene(ctx, model_path(name))
    return sc, ray.SceneDescription.from_scene(ctx, sc)


# ---- acceleration structure ---------------------------------------------------------------------------

def _walk_accel(nodes, tris):
    """CPU walk of the 8-wide tree: returns per-triangle visit counts and checks box containment."""
    n_tris = tris.shape[0]
    seen = np.zeros(n_tris, dtype=np.int64)
    f = nodes.view(np.float32)

    def child_box(node, i):
        w = nodes[node]
        e = int(w[3])
        s = [np.float32(2.0) ** np.float32(((e >> (8 * a)) & 0xFF) - 127) for a in range(3)]
        word, sh = i >> 2, 8 * (i & 3)
        lo = [f[node][a] + np.float32((int(w[8 + 2 * a + word]) >> sh) & 0xFF) * s[a] for a in range(3)]
        hi = [f[node][a] + np.float32((int(w[14 + 2 * a + word]) >> sh) & 0xFF) * s[a] for a in range(3)]
        return np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)

    max_depth = 0
    stack = [(0, None, None, 1)]
    n_nodes_seen = 0
    while stack:
        node, plo, phi, depth = stack.pop()
        n_nodes_seen += 1
        max_depth = max(max_depth, depth)
        w = nodes[node]
        imask = int(w[3]) >> 24
        child_base, tri_base, count0 = int(w[4]), int(w[5]) & 0x0FFFFFFF, int(w[5]) >> 28
        for i in range(8):
            tmask = (int(w[6 + (i >> 2)]) >> (8 * (i & 3))) & 0xFF  # bvh.cuh: a leaf's contribution to the triangle hit mask
            inner = (imask >> i) & 1
            if not inner and tmask == 0:
                continue
            lo, hi = child_box(node, i)
            if plo is not None:
                # parent's box for this node >= the node's own child boxes: both are conservative roundings of the same true
                # bounds, so a child box may poke out of the parent's box by at most the node's own quantisation cells
                cell = np.array([2.0 ** (((int(w[3]) >> (8 * a)) & 0xFF) - 127) for a in range(3)], dtype=np.float64)
                slack = 2.0 * cell + 1e-6 * (1.0 + np.abs(plo) + np.abs(phi))
                assert np.all(lo >= plo - slack) and np.all(hi <= phi + slack), \
                    "child box of node %d slot %d outside its parent's box: %s %s vs %s %s" % (node, i, lo, hi, plo, phi)
                assert np.all(hi >= lo)
            if inner:
                assert tmask == 0
                rel = bin(imask & ((1 << i) - 1)).count("1")
                stack.append((child_base + rel, lo, hi, depth + 1))
            else:
                cnt = bin(tmask).count("1")
                first = (tmask & -tmask).bit_length() - 1
                assert cnt in (1, 2) and tmask == ((1 << cnt) - 1) << first, "a leaf owns a run of 1 or 2 triangle bits"
                off = first + (0 if i < 4 else count0)
                for k in range(cnt):
                    t = tri_base + off + k
                    seen[t] += 1
                    v = tris[t].reshape(3, 4)[:, :3].astype(np.float64)
                    assert np.all(v >= lo - 1e-6 * (1 + np.abs(lo))) and np.all(v <= hi + 1e-6 * (1 + np.abs(hi))), \
                        "triangle outside its leaf box"
    return seen, n_nodes_seen, max_depth


@pytest.mark.parametrize("name", ["cornell", "Duck", "tunnel"])
def test_accel_invariants(sol, ctx, name):
    fs, osc = oracle_scene(name)
    sc, sd = _product(sol, ctx, name)
    info = sd.accel_info()
    assert info.n_instances == len(fs.instances) and info.n_triangles == osc.tri_count
    nodes, tris = sd.read_nodes(), sd.read_triangles()
    seen, n_nodes, depth = _walk_accel(nodes, tris)
    assert np.all(seen == 1), "every triangle must be reachable exactly once"
    assert n_nodes == info.n_wide_nodes and depth == info.wide_depth
    # ids carried in the w lanes: (instance, primitive, global ordinal) cover the scene exactly once
    ids = tris.view(np.uint32).reshape(-1, 3, 4)[:, :, 3]
    first = np.cumsum([0] + [i["n_indices"] // 3 for i in fs.instances])
    assert sorted(ids[:, 2].tolist()) == list(range(osc.tri_count))
    assert np.array_equal(ids[:, 2], first[ids[:, 0]] + ids[:, 1])
    lo, hi = osc.bounds()
    np.testing.assert_allclose(np.array(info.scene_lo[:]), lo, atol=1e-5)
    np.testing.assert_allclose(np.array(info.scene_hi[:]), hi, atol=1e-5)
    assert 0 < info.sah_cost_binary <= info.sah_cost_lbvh * 1.0001  # treelet pass never makes SAH worse
    # the world-space vertices equal transform * position in f32 up to rounding
    g = int(ids[0, 2])
    inst = int(ids[0, 0])
    I = fs.instances[inst]
    idx = fs.indices[I["first_index"] + 3 * int(ids[0, 1]): I["first_index"] + 3 * int(ids[0, 1]) + 3] + I["first_vertex"]
    P = fs.vertices[idx, 0:3].astype(np.float64) @ I["transform"][:3, :3].astype(np.float64) + I["transform"][3, :3]
    np.testing.assert_allclose(tris[0].reshape(3, 4)[:, :3], P, rtol=1e-5, atol=1e-6)
    assert g < osc.tri_count


# ---- primary-hit ids (north_star gate 1) ---------------------------------------------------------------

@pytest.mark.parametrize("name,w,h", [("Duck", 900, 600), ("cornell", 512, 512), ("tunnel", 1920, 1080)])
def test_primary_hit_ids(sol, ctx, name, w, h):
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    fs, osc = oracle_scene(name)
    ou = ocam.scene_uniforms(oracle_camera(fs, name, w, h), w, h, 0)
    o_rgba, o_ids, o_bt, o_flags = osc.debug(ou, w, h)
    sc, sd = _product(sol, ctx, name)
    cam = product_camera(sc, name, w, h)
    u = scene.scene_uniforms(cam, w, h, 0)
    assert bytes(u) == ou, "product and oracle uniform blocks must be identical inputs"
    render = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
    ids = sol.Image2d(ctx, w, h, N.FORMAT_RG32UI)
    sbt = simple_pipeline(ctx, "debug")
    sbt.cmd_trace_rays(ray.TraceBindings(sd, u, None, render, ids), (w, h, 1))
    g_ids, g_rgba = ids.readback(), render.readback()
    mism = np.any(g_ids != o_ids, axis=2)
    listed = o_flags != 0  # the oracle's edge / tie / near-miss list
    assert (mism & ~listed).sum() == 0, "hit ids differ on %d unlisted pixels" % (mism & ~listed).sum()
    assert listed.mean() < 0.005
    # debug.rchit colours: barycentrics (hit) or direction (miss) within 1 LSB away from listed pixels
    d = np.abs(g_rgba.astype(np.int32) - o_rgba.astype(np.int32)).max(axis=2)
    assert (d[~listed] > 1).sum() == 0
    assert np.all(g_rgba[..., 3] == 0)


def test_random_rays_vs_oracle(sol, ctx):
    """traceRayEXT on 10^6 incoherent rays: (instance, primitive) equal to the f64 oracle except listed edge/tie rays."""
    fs, osc = oracle_scene("tunnel")
    sc, sd = _product(sol, ctx, "tunnel")
    rng = np.random.default_rng(11)
    n = 1_000_000
    lo, hi = osc.bounds()
    o = rng.uniform(lo * 0.9, hi * 0.9, size=(n, 3))
    d = rng.normal(size=(n, 3))
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)
    g_hits, g_t = sd.trace_rays(rays)
    o_hits, o_t, flags = osc.trace_rays(rays, classify=True)
    mism = np.any(g_hits[:, :2] != o_hits[:, :2], axis=1)
    assert (mism & (flags == 0)).sum() == 0
    assert (flags != 0).mean() < 0.01
    ok = ~mism & (o_hits[:, 0] != oracle.MISS)
    # t = plane equation through the stored f32 world-space vertices: rounding a vertex of tunnel.gltf's
    # 9.5 x 0.05 sliver triangles by half an ulp tilts their plane by ~5e-6 rad, i.e. up to ~1e-4 along the
    # long side.  t is never read by the reference's shaders; it only orders hits.
    np.testing.assert_allclose(g_t[ok], o_t[ok], rtol=1e-4, atol=5e-4)
    assert np.median(np.abs(g_t[ok] - o_t[ok]) / o_t[ok]) < 2e-6
    gu = g_hits[ok, 2:].view(np.float32)
    ou = o_hits[ok, 2:].view(np.float32)
    du = np.abs(gu - ou).max(axis=1)  # barycentrics are ill-conditioned on slivers hit at grazing angles
    assert np.quantile(du, 0.999) < 1e-4 and du.max() < 5e-2


# ---- path tracing ----------------------------------------------------------------------------------------

def _render_gpu(sol, ctx, name, w, h, frames, sky, spp, mb, schedule, accum_mode=0, start=0, collect=False, two_level=False):
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    sc, sd = _product_two_level(sol, ctx, name) if two_level else _product(sol, ctx, name)
    cam = product_camera(sc, name, w, h)
    accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    render = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
    sbt = pathtrace_pipeline(ctx, sky)
    for f in frames:
        u = scene.scene_uniforms(cam, w, h, f)
        sd.tlas_regenerate()
        sbt.cmd_trace_rays(ray.TraceBindings(sd, u, accum, render, accumulation_start_frame=start, samples_per_frame=spp,
                                             max_bounces=mb, schedule=schedule, accum_mode=accum_mode, collect_stats=collect),
                           (w, h, 1))
    return accum.readback(), render.readback()


def _render_oracle(name, w, h, frames, sky, spp, mb, start=0):
    fs, osc = oracle_scene(name)
    cam = oracle_camera(fs, name, w, h)
    acc = np.zeros((h, w, 4), dtype=np.float32)
    st = oracle.OrcStats()
    rgba = None
    for f in frames:
        rgba, _ = osc.pathtrace_frame(ocam.scene_uniforms(cam, w, h, f), w, h, acc, start, sky, spp, mb, st)
    return acc, rgba, st


@pytest.mark.parametrize("schedule", [0, 1, 3])
@pytest.mark.parametrize("name,w,h,sky,mb", [("cornell", 128, 128, False, 32), ("cornell", 96, 64, False, 4),
                                             ("tunnel", 160, 90, True, 32), ("tunnel", 160, 90, True, 8)])
def test_pathtrace_frames_vs_oracle(sol, ctx, name, w, h, sky, mb, schedule):
    """Identical per-pixel RNG streams: frames 0..1 agree with the oracle except decision-flip pixels."""
    ctx.reset_stats()
    g_acc, g_rgba = _render_gpu(sol, ctx, name, w, h, [0, 1], sky, 8, mb, schedule)
    o_acc, o_rgba, st = _render_oracle(name, w, h, [0, 1], sky, 8, mb)
    gs = ctx.stats()
    assert gs.paths == st.paths == w * h * 8 * 2
    assert abs(int(gs.rays) - int(st.rays)) <= 0.002 * st.rays  # rays/path statistics agree
    d = np.abs(g_acc[..., :3] - o_acc[..., :3])
    differing = (d.max(axis=2) > 1e-3 * (1.0 + np.abs(o_acc[..., :3]).max(axis=2))).mean()
    assert differing < 0.02, "too many pixels differ from the oracle: %.4f" % differing
    mre, psnr = image_metrics(g_acc, o_acc)
    assert mre < 0.01
    assert np.all(g_acc[..., 3] == 1.0)
    assert (np.abs(g_rgba.astype(np.int32) - o_rgba.astype(np.int32)).max(axis=2) > 1).mean() < 0.02


def test_wavefront_equals_megakernel(sol, ctx):
    a, _ = _render_gpu(sol, ctx, "tunnel", 256, 144, [0, 1, 2], True, 8, 32, 0)
    b, _ = _render_gpu(sol, ctx, "tunnel", 256, 144, [0, 1, 2], True, 8, 32, 1)
    d = np.abs(a - b)[..., :3]
    # same per-ray functions, but the two kernels are compiled separately (FMA contraction differs), so a
    # few decision-flip pixels are expected
    assert (d.max(axis=2) > 1e-4 * (1 + np.abs(b[..., :3]).max(axis=2))).mean() < 0.01
    assert d.sum() / b[..., :3].sum() < 2e-3


@pytest.mark.parametrize("name,w,h,sky,mb,two_level", [("tunnel", 256, 144, True, 8, False), ("tunnel", 1920, 1080, True, 8, False),
                                                       ("cornell", 200, 120, False, 32, False), ("Duck", 160, 120, True, 8, True),
                                                       ("tunnel", 192, 108, True, 8, True)])
def test_warpfront_equals_wavefront(sol, ctx, name, w, h, sky, mb, two_level):
    """The warp-local wavefront kernel (schedule 3) runs the per-ray / per-pixel functions of the queue-based wavefront
    (k_wf_generate / k_wf_trace / k_wf_shade / k_wf_resolve) in a different order only: same rays, same paths, and the same
    image up to decision-flip pixels (separately compiled instantiations: FMA contraction and the order of equal-t tests)."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    # ONE scene description for all renders: a rebuilt hierarchy may number its nodes differently (atomics in the builder),
    # which changes the order equal-t candidates are tested in
    sc, sd = _product_two_level(sol, ctx, name) if two_level else _product(sol, ctx, name)
    cam = product_camera(sc, name, w, h)
    sbt = pathtrace_pipeline(ctx, sky)

    def render(schedule):
        accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
        rend = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
        ctx.reset_stats()
        for f in range(3):
            sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, rend, samples_per_frame=8,
                                                 max_bounces=mb, schedule=schedule), (w, h, 1))
        return accum.readback(), rend.readback(), ctx.stats()

    a, ra, s0 = render(N.SCHEDULE_WAVEFRONT)
    b, rb, s1 = render(N.SCHEDULE_WARPFRONT)
    assert s0.paths == s1.paths == 3 * 8 * w * h
    assert abs(int(s0.rays) - int(s1.rays)) <= 1e-4 * int(s0.rays) + 2
    d = np.abs(a - b)[..., :3]
    assert (d.max(axis=2) > 1e-4 * (1 + np.abs(b[..., :3]).max(axis=2))).mean() < 0.01
    assert d.sum() / b[..., :3].sum() < 2e-3
    assert np.all(b[..., 3] == 1.0) and np.all(np.isfinite(b))
    # and it is a pure function of its inputs: a second run gives the same bits although slot order and atomics vary
    c, rc, s2 = render(N.SCHEDULE_WARPFRONT)
    assert int(s2.rays) == int(s1.rays)
    assert np.array_equal(b, c) and np.array_equal(rb, rc)


def test_accumulation_restart_and_alpha(sol, ctx):
    """pathtrace.rgen:89-101: alpha = 1/(frame+1-start); a restart overwrites whatever the image held."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    w = h = 64
    sc, sd = _product(sol, ctx, "cornell")
    cam = product_camera(sc, "cornell", w, h)
    sbt = pathtrace_pipeline(ctx, False)
    accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    accum.upload(np.full((h, w, 4), 77.0, dtype=np.float32))
    sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, 5), accum, None, accumulation_start_frame=5), (w, h, 1))
    a5 = accum.readback()
    fresh = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, 5), fresh, None, accumulation_start_frame=5), (w, h, 1))
    np.testing.assert_array_equal(a5, fresh.readback())
    sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, 6), accum, None, accumulation_start_frame=5), (w, h, 1))
    a56 = accum.readback()
    only6 = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, 6), only6, None, accumulation_start_frame=6), (w, h, 1))
    np.testing.assert_allclose(a56[..., :3], 0.5 * a5[..., :3] + 0.5 * only6.readback()[..., :3], rtol=1e-5, atol=1e-6)


def test_sum_mode_matches_running_mean(sol, ctx):
    """per-rank sums + resolve equal the reference's running mix up to fp rounding."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray

    mix, _ = _render_gpu(sol, ctx, "cornell", 64, 64, range(6), False, 8, 8, 0)
    s, _ = _render_gpu(sol, ctx, "cornell", 64, 64, range(6), False, 8, 8, 0, accum_mode=N.ACCUM_SUM)
    assert np.all(s[..., 3] == 6.0)
    np.testing.assert_allclose(s[..., :3] / 6.0, mix[..., :3], rtol=2e-5, atol=1e-6)
    tgt = sol.Image2d(ctx, 64, 64, N.FORMAT_RGBA32F)
    tgt.upload(s)
    out = sol.Image2d(ctx, 64, 64, N.FORMAT_RGBA32F)
    rgba = sol.Image2d(ctx, 64, 64, N.FORMAT_RGBA8)
    ray.resolve_sum(ctx, tgt, out, rgba)
    np.testing.assert_allclose(out.readback()[..., :3], mix[..., :3], rtol=2e-5, atol=1e-6)
    assert np.all(rgba.readback()[..., 3] == 255)


def test_instance_transform_update(sol, ctx):
    """SceneDescription::blas_transform + tlas_regenerate: hits follow the new transform (oracle rebuilt with it)."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    w = h = 128
    fs, _ = oracle_scene("cornell")
    sc, sd = _product(sol, ctx, "cornell")
    T = np.eye(4, dtype=np.float32)
    T[3, :3] = [0.3, 0.2, -0.1]  # [col][row] storage: translation column
    new_t = oracle.gltf_flatten.mat4_mul(T, fs.instances[6]["transform"])
    transforms = [i["transform"] for i in fs.instances]
    transforms[6] = new_t
    osc = oracle.Scene(fs, transforms)
    sd.blas_transform(new_t.reshape(16), 6)
    sd.update()
    sd.tlas_regenerate()
    inst = sd.instances()[6]
    np.testing.assert_allclose(np.array(inst.transform[:]), new_t.reshape(16), rtol=0, atol=0)
    tit = oracle.gltf_flatten.mat4_inverse(new_t).T.reshape(16)
    np.testing.assert_allclose(np.array(inst.transform_it[:]), tit, rtol=1e-6, atol=1e-7)
    cam = product_camera(sc, "cornell", w, h)
    u = scene.scene_uniforms(cam, w, h, 0)
    ids = sol.Image2d(ctx, w, h, N.FORMAT_RG32UI)
    simple_pipeline(ctx, "debug").cmd_trace_rays(ray.TraceBindings(sd, u, None, None, ids), (w, h, 1))
    _, o_ids, _, flags = osc.debug(bytes(u), w, h)
    mism = np.any(ids.readback() != o_ids, axis=2)
    assert (mism & (flags == 0)).sum() == 0
    # and shading uses the updated transform too
    a, _ = None, None
    accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    pathtrace_pipeline(ctx, False).cmd_trace_rays(ray.TraceBindings(sd, u, accum, None), (w, h, 1))
    ref = np.zeros((h, w, 4), dtype=np.float32)
    osc.pathtrace_frame(bytes(u), w, h, ref)
    d = np.abs(accum.readback()[..., :3] - ref[..., :3])
    assert (d.max(axis=2) > 1e-3 * (1 + ref[..., :3].max(axis=2))).mean() < 0.02


def test_ao_frame_vs_oracle(sol, ctx):
    """4-ray-ao on Duck (ToyCar.glb is missing upstream) with the example's camera."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    w, h = 240, 135
    blue = load_blue_noise()
    fs, osc = oracle_scene("Duck")
    sc, sd = _product(sol, ctx, "Duck")
    ctx.set_blue_noise(blue)
    cam = product_camera(sc, "Duck_ao", w, h)
    img = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    ref = np.zeros((h, w, 4), dtype=np.float32)
    sbt = simple_pipeline(ctx, "ao")
    ocamera = oracle_camera(fs, "Duck_ao", w, h)
    for f in range(2):
        u = scene.scene_uniforms(cam, w, h, f)
        assert bytes(u) == ocam.scene_uniforms(ocamera, w, h, f)
        sbt.cmd_trace_rays(ray.TraceBindings(sd, u, img, None), (w, h, 1))
        osc.ao_frame(bytes(u), w, h, ref, blue)
    g = img.readback()
    d = np.abs(g[..., :3] - ref[..., :3]).max(axis=2)
    assert (d > 1e-4).mean() < 0.01
    assert np.all(g[..., 3] == 1.0) and g[..., :3].min() >= 0.0 and g[..., :3].max() <= 1.0


def test_stats_counters_and_roofline_inputs(sol, ctx):
    """collect_stats: nodes / triangles per ray are what the B_ray roofline model is computed from."""
    ctx.reset_stats()
    _render_gpu(sol, ctx, "tunnel", 160, 90, [0], True, 8, 8, 0, collect=True)
    st = ctx.stats()
    assert st.rays > 0 and st.nodes_visited > st.rays and st.tris_tested > 0
    assert 1.0 < st.nodes_visited / st.rays < 200.0


def test_error_paths(sol, ctx):
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    L = N.lib()
    assert L.solb_ctx_create(0, None, None) == -1
    bad = ctypes.c_void_p()
    assert L.solb_ctx_create(9999, None, ctypes.byref(bad)) == -1
    sc, sd = _product(sol, ctx, "cornell")
    cam = product_camera(sc, "cornell", 32, 32)
    u = scene.scene_uniforms(cam, 32, 32, 0)
    a = sol.Image2d(ctx, 32, 32, N.FORMAT_RGBA32F)
    r_wrong = sol.Image2d(ctx, 16, 16, N.FORMAT_RGBA8)
    sbt = pathtrace_pipeline(ctx, False)
    with pytest.raises(sol.SolbError):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, u, a, r_wrong), (32, 32, 1))  # size mismatch
    with pytest.raises(sol.SolbError):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, u, a, None), (64, 32, 1))  # extent mismatch
    with pytest.raises(sol.SolbError):
        sol.Image2d(ctx, 0, 4, N.FORMAT_RGBA32F)
    with pytest.raises(sol.SolbError):
        sd.blas_transform(np.eye(4, dtype=np.float32), 99)
    # a section without indices is rejected like the reference's RT path cannot represent it (App.A item 4)
    m = sc.meshes[0]
    m2 = scene.Mesh("x", m.vertices, m.indices, m.transform, [scene.PrimitiveSection(0, 0, m.vertices.shape[0], 0, 0, 0)])
    with pytest.raises(sol.SolbError):
        ray.SceneDescription.from_meshes(ctx, [m2], [m.transform], sc.materials)
    # empty scene: builds, every ray misses
    empty = ray.SceneDescription.from_meshes(ctx, [], [], np.zeros((0, 12), np.float32))
    hits, _ = empty.trace_rays(np.array([[0, 0, 0, 0, 0, 0, 1, 10]], dtype=np.float32))
    assert hits[0, 0] == N.MISS


# ---- north_star gate 2: converged images vs the committed 4096-spp golden fixtures ---------------------

import os as _os

GOLDEN = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,w,h,sky,mb", [("tunnel", 160, 90, True, 8), ("cornell", 96, 96, False, 32)])
@pytest.mark.parametrize("schedule,two_level", [(0, False), (1, False), (0, True), (3, False), (3, True)])
def test_converged_image_vs_golden(sol, ctx, name, w, h, sky, mb, schedule, two_level):
    """512 frames x 8 spp = 4096 spp: mean relative error < 1 % and PSNR > 40 dB (BASELINE.json north_star)."""
    g = np.load(_os.path.join(GOLDEN, "converged_%s_%dx%d_4096spp_b%d.npz" % (name, w, h, mb)))
    ctx.reset_stats()
    acc, rgba = _render_gpu(sol, ctx, name, w, h, range(512), sky, 8, mb, schedule, two_level=two_level)
    st = ctx.stats()
    mre, psnr = image_metrics(acc, g["accum"])
    assert mre < 0.01, "mean relative error %.4f" % mre
    assert psnr > 40.0, "PSNR %.1f dB" % psnr
    # the display image too (gamma 2.2 rgba8 as stored by the oracle)
    d = rgba[..., :3].astype(np.float64) - g["rgba8"][..., :3].astype(np.float64)
    assert 10 * np.log10(255.0 ** 2 / max(np.mean(d ** 2), 1e-12)) > 40.0
    # structural statistics: rays per path equal to the oracle's within 0.1 %
    assert st.paths == int(g["paths"])
    assert abs(st.rays / st.paths - int(g["rays"]) / int(g["paths"])) < 1e-3 * int(g["rays"]) / int(g["paths"])


@pytest.mark.parametrize("name,cam_name,w,h", [("cornell", "cornell", 256, 256), ("Duck", "Duck", 450, 300), ("tunnel", "tunnel", 480, 270)])
def test_primary_hit_ids_vs_golden(sol, ctx, name, cam_name, w, h):
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    g = np.load(_os.path.join(GOLDEN, "hitids_%s_%dx%d.npz" % (name, w, h)))
    sc, sd = _product(sol, ctx, name)
    cam = product_camera(sc, cam_name, w, h)
    ids = sol.Image2d(ctx, w, h, N.FORMAT_RG32UI)
    simple_pipeline(ctx, "debug").cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, 0), None, None, ids), (w, h, 1))
    got = ids.readback()
    inst = np.where(got[..., 0] == N.MISS, 255, got[..., 0]).astype(np.uint8)
    prim = np.where(got[..., 1] == N.MISS, 65535, got[..., 1]).astype(np.uint16)
    mism = (inst != g["inst"]) | (prim != g["prim"])
    assert (mism & (g["flags"] == 0)).sum() == 0


@pytest.mark.parametrize("accel", ["flat", "two_level"])
def test_ao_vs_golden(sol, ctx, accel):
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    g = np.load(_os.path.join(GOLDEN, "ao_duck_160x90_f0-3.npz"))
    sc = scene.load_scene(ctx, model_path("Duck"))
    sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=N.ACCEL_TWO_LEVEL if accel == "two_level" else N.ACCEL_FLAT)
    ctx.set_blue_noise(load_blue_noise())
    cam = product_camera(sc, "Duck_ao", 160, 90)
    img = sol.Image2d(ctx, 160, 90, N.FORMAT_RGBA32F)
    sbt = simple_pipeline(ctx, "ao")
    ctx.reset_stats()
    for f in range(4):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, 160, 90, f), img, None), (160, 90, 1))
    d = np.abs(img.readback()[..., :3] - g["image"]).max(axis=2)
    assert (d > 1e-4).mean() < 0.01
    st = ctx.stats()
    assert st.paths == int(g["paths"]) and abs(int(st.rays) - int(g["rays"])) <= 0.002 * int(g["rays"])


def test_cpp_host_example_runs_the_reference_call_sequence(sol, ctx):
    """examples/pathtrace_offscreen.cpp = examples/5-pathtrace.rs through the C++ mirror (sol::ray::*): its final
    frame must be byte-identical to the same frames rendered through the Python wrapper (same C ABI below)."""
    import json
    import subprocess

    from helpers import ROOT
    from sol_rs_b200 import _native as N

    exe = _os.path.join(ROOT, "examples", "pathtrace_offscreen")
    if not _os.path.exists(exe):  # a build artefact (git-ignored): link it against the in-tree libraries when it did not travel
        subprocess.run(["make", "-C", _os.path.join(ROOT, "sol_rs_b200", "csrc"), "../../examples/pathtrace_offscreen"], check=False,
                       capture_output=True, timeout=300)
    assert _os.path.exists(exe), "run __graft_entry__.build()"
    out = subprocess.run([exe, "--model", "models/cornell.gltf", "--frames", "3", "--size", "160x120"], capture_output=True,
                         text=True, cwd=ROOT, timeout=120)
    assert out.returncode == 0, out.stderr
    info = json.loads(out.stdout.strip().splitlines()[-1])
    _, rgba = _render_gpu(sol, ctx, "cornell", 160, 120, range(3), False, 8, 32, N.SCHEDULE_AUTO)  # the example keeps the default schedule
    h = 1469598103934665603
    for b in rgba.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert info["fnv1a"] == "%016x" % h
    assert info["rays"] > 0
    # the same frames through the two-level structure (C++ mirror: set_accel_mode + accel_build)
    out2 = subprocess.run([exe, "--model", "models/cornell.gltf", "--frames", "3", "--size", "160x120", "--two-level"], capture_output=True,
                          text=True, cwd=ROOT, timeout=120)
    assert out2.returncode == 0, out2.stderr
    info2 = json.loads(out2.stdout.strip().splitlines()[-1])
    assert info2["rays"] > 0 and abs(info2["rays"] - info["rays"]) <= 0.002 * info["rays"]
    # error behaviour: the reference panics without --model
    bad = subprocess.run([exe], capture_output=True, text=True, cwd=ROOT, timeout=60)
    assert bad.returncode != 0 and "no gltf file given" in bad.stderr


@pytest.mark.parametrize("n,bits,kind", [(1, 63, "random"), (31, 63, "random"), (2048, 63, "random"), (2049, 63, "random"),
                                         (100_000, 63, "random"), (1_000_003, 63, "random"), (300_000, 63, "equal"),
                                         (300_000, 63, "sorted"), (300_000, 63, "reversed"), (300_000, 16, "fewbits"),
                                         (5_000_000, 63, "random")])
def test_onesweep_sort(sol, ctx, n, bits, kind):
    """The builder's onesweep radix sort: sorted by key, stable (values of equal keys keep their input order)."""
    from sol_rs_b200 import _native as N

    rng = np.random.default_rng(n + bits)
    if kind == "equal":
        keys = np.full(n, 0x1234567890ABCDEF & ((1 << 63) - 1), dtype=np.uint64)
    elif kind == "sorted":
        keys = np.sort(rng.integers(0, 1 << 63, size=n, dtype=np.uint64))
    elif kind == "reversed":
        keys = np.sort(rng.integers(0, 1 << 63, size=n, dtype=np.uint64))[::-1].copy()
    elif kind == "fewbits":
        keys = rng.integers(0, 1 << 10, size=n, dtype=np.uint64)  # many duplicates
    else:
        keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32)
    k, v = keys.copy(), vals.copy()
    N.check(N.lib().solb_test_sort_pairs(ctx.handle, k.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p), n, bits),
            ctx.handle)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


def test_accumulation_checkpoint_resume(sol, ctx, tmp_path):
    """frames 0..5 in one go == frames 0..2, checkpoint, restore, frames 3..5 (bit-identical)."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import io, ray, scene

    w, h = 96, 64
    sc, sd = _product(sol, ctx, "cornell")
    cam = product_camera(sc, "cornell", w, h)
    sbt = pathtrace_pipeline(ctx, False)

    def run(accum, frames):
        for f in frames:
            sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, None, max_bounces=8), (w, h, 1))

    full = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    run(full, range(6))
    part = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    run(part, range(3))
    io.save_checkpoint(str(tmp_path / "ck"), part, 0, 3)
    restored, start, nxt = io.load_checkpoint(str(tmp_path / "ck"), ctx)
    assert (start, nxt) == (0, 3)
    run(restored, range(nxt, 6))
    np.testing.assert_array_equal(restored.readback(), full.readback())
    render = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
    sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, 6), restored, render, max_bounces=8), (w, h, 1))
    io.write_png(str(tmp_path / "frame.png"), render.readback())


# ---- two-level mode: TLAS over shared object-space BLASes -----------------------------

def _product_two_level(sol, ctx, name):
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    sc = scene.load_scene(ctx, model_path(name))
    return sc, ray.SceneDescription.from_scene(ctx, sc, accel_mode=N.ACCEL_TWO_LEVEL)


@pytest.mark.parametrize("name,w,h", [("Duck", 900, 600), ("cornell", 512, 512), ("tunnel", 960, 540)])
def test_two_level_primary_hit_ids(sol, ctx, name, w, h):
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    fs, osc = oracle_scene(name)
    ou = ocam.scene_uniforms(oracle_camera(fs, name, w, h), w, h, 0)
    _, o_ids, _, o_flags = osc.debug(ou, w, h)
    sc, sd = _product_two_level(sol, ctx, name)
    info = sd.accel_info()
    assert info.mode == N.ACCEL_TWO_LEVEL and info.n_blas == len(fs.instances) and info.n_tlas_nodes >= 1
    assert info.n_triangles == osc.tri_count and info.tlas_depth >= 1 and info.wide_depth > info.tlas_depth
    u = scene.scene_uniforms(product_camera(sc, name, w, h), w, h, 0)
    ids = sol.Image2d(ctx, w, h, N.FORMAT_RG32UI)
    render = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
    simple_pipeline(ctx, "debug").cmd_trace_rays(ray.TraceBindings(sd, u, None, render, ids), (w, h, 1))
    mism = np.any(ids.readback() != o_ids, axis=2)
    assert (mism & (o_flags == 0)).sum() == 0, "hit ids differ on %d unlisted pixels" % (mism & (o_flags == 0)).sum()
    # object-space triangles: the stored vertices are the file's, bit for bit
    tris = sd.read_triangles()
    tid = tris.view(np.uint32).reshape(-1, 3, 4)[:, :, 3]
    assert sorted(tid[:, 2].tolist()) == list(range(osc.tri_count))
    k = int(np.argmin(tid[:, 2]))
    I = fs.instances[0]
    idx = fs.indices[I["first_index"] + 3 * int(tid[k, 1]): I["first_index"] + 3 * int(tid[k, 1]) + 3] + I["first_vertex"]
    assert np.array_equal(tris[k].reshape(3, 4)[:, :3], fs.vertices[idx, 0:3])


def _instanced_duck(sol, ctx, mode):
    from helpers import duck_extras, instanced_variant
    from sol_rs_b200 import ray, scene

    fs, _ = oracle_scene("Duck")
    extras = duck_extras(fs)
    sc = scene.load_scene(ctx, model_path("Duck"))
    sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=mode)
    for src, t, mat in extras:
        sd.add_instance(src, t, mat)
    sd.accel_build()
    return fs, extras, oracle.Scene(instanced_variant(fs, extras)), sc, sd


def _sphere_rays(osc, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = osc.bounds()
    c, r = (lo + hi) / 2, np.linalg.norm(hi - lo)
    o = c + rng.normal(size=(n, 3)) * r
    d = (c + rng.normal(size=(n, 3)) * 0.2 * r) - o
    return np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)


def test_two_level_shared_blas_instances(sol, ctx):
    """four instances of ONE BLAS (mirrored, scaled, overlapping): two-level hits = oracle = flattened build, and the
    two-level structure stores the geometry once."""
    from sol_rs_b200 import _native as N

    results = {}
    for mode in (N.ACCEL_TWO_LEVEL, N.ACCEL_FLAT):
        fs, extras, osc, sc, sd = _instanced_duck(sol, ctx, mode)
        rays = _sphere_rays(osc, 400_000, 7)
        o_hits, o_t, flags = osc.trace_rays(rays, classify=True)
        hits, t = sd.trace_rays(rays)
        mism = np.any(hits[:, :2] != o_hits[:, :2], axis=1)
        assert (mism & (flags == 0)).sum() == 0
        ok = (o_hits[:, 0] != oracle.MISS) & ~mism
        np.testing.assert_allclose(t[ok], o_t[ok], rtol=2e-4, atol=2e-5)
        assert set(np.unique(o_hits[:, 0]).tolist()) >= {0, 1, 2, 3}
        results[mode] = sd.accel_info()
        assert [i.id for i in sd.instances()] == [0, 1, 2, 3]
    two, flat = results[N.ACCEL_TWO_LEVEL], results[N.ACCEL_FLAT]
    assert two.n_instances == 4 and two.n_blas == 1 and two.n_triangles == 4212
    assert flat.n_triangles == 4 * 4212 and flat.n_wide_nodes > 3 * (two.n_wide_nodes - two.n_instances)


@pytest.mark.parametrize("schedule", ["wavefront", "megakernel", "warpfront"])
@pytest.mark.parametrize("name,w,h,sky,mb", [("cornell", 128, 128, False, 32), ("tunnel", 192, 108, True, 8)])
def test_two_level_pathtrace_matches_oracle(sol, ctx, name, w, h, sky, mb, schedule):
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    sc, sd = _product_two_level(sol, ctx, name)
    cam = product_camera(sc, name, w, h)
    accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    sbt = pathtrace_pipeline(ctx, sky)
    ctx.reset_stats()
    for f in range(2):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, None, samples_per_frame=8, max_bounces=mb,
                                             schedule={"wavefront": N.SCHEDULE_WAVEFRONT, "megakernel": N.SCHEDULE_MEGAKERNEL,
                                                       "warpfront": N.SCHEDULE_WARPFRONT}[schedule],
                                             collect_stats=True), (w, h, 1))
    st = ctx.stats()
    o_acc, _, o_st = _render_oracle(name, w, h, range(2), sky, 8, mb)
    a, b = accum.readback()[..., :3], o_acc[..., :3]
    d = np.abs(a - b)
    assert (d.max(axis=2) > 1e-3 * (1 + b.max(axis=2))).mean() < 0.02
    assert d.sum() / b.sum() < 0.01
    assert abs(int(st.rays) - int(o_st.rays)) <= 0.002 * int(o_st.rays) and st.nodes_visited > st.rays


def test_two_level_tlas_regenerate_moves_instances(sol, ctx):
    """blas_transform + tlas_regenerate in two-level mode rebuilds only the TLAS (SceneDescription::tlas_regenerate,
    src/ray/mod.rs:162-193): hits follow the moved instance, BLAS nodes and triangles stay byte-identical."""
    from helpers import instanced_variant, trs
    from sol_rs_b200 import _native as N

    fs, extras, _, sc, sd = _instanced_duck(sol, ctx, N.ACCEL_TWO_LEVEL)
    info0 = sd.accel_info()
    nodes0, tris0 = sd.read_nodes(), sd.read_triangles()
    moved = np.ascontiguousarray((trs((0.5, 1.0, -2.0), (0, 1, 0), 2.0).T.astype(np.float64)
                                  @ np.asarray(extras[1][1], dtype=np.float64).T).T, dtype=np.float32)
    sd.blas_transform(moved, 2)
    sd.tlas_regenerate()
    info1 = sd.accel_info()
    nodes1, tris1 = sd.read_nodes(), sd.read_triangles()
    cap = info0.n_instances
    assert info1.n_wide_nodes == info0.n_wide_nodes and np.array_equal(nodes0[cap:], nodes1[cap:]) and np.array_equal(tris0, tris1)
    assert not np.array_equal(nodes0[:cap], nodes1[:cap])
    osc = oracle.Scene(instanced_variant(fs, [extras[0], (0, moved, 0), extras[2]]))
    rays = _sphere_rays(osc, 200_000, 13)
    o_hits, _, flags = osc.trace_rays(rays, classify=True)
    hits, _ = sd.trace_rays(rays)
    assert (np.any(hits[:, :2] != o_hits[:, :2], axis=1) & (flags == 0)).sum() == 0
    assert (o_hits[:, 0] == 2).sum() > 500
    # the reference's transform_it follows the transform (SceneInstance::update_transform, src/ray/mod.rs:27-30)
    inst = sd.instances()[2]
    np.testing.assert_allclose(np.array(inst.transform[:]).reshape(4, 4), moved, rtol=0, atol=0)
    sd.tlas_regenerate()  # clean: no-op
    assert sd.accel_info().n_wide_nodes == info1.n_wide_nodes


def test_two_level_many_instances_tlas_rebuild_time(sol, ctx):
    """1 000 instances of two shared BLASes (config 5's instance count): build, per-frame TLAS regenerate, hit parity
    against the flattened build of the same instances."""
    from helpers import trs
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    rng = np.random.default_rng(3)
    sc = scene.load_scene(ctx, model_path("cornell"))
    xf = [trs(rng.uniform(-20, 20, 3), rng.normal(size=3), rng.uniform(0, 6.28), (rng.uniform(0.5, 1.5),) * 3) for _ in range(992)]
    sds = {}
    for key, mode, fast in (("two", N.ACCEL_TWO_LEVEL, "1"), ("two_general", N.ACCEL_TWO_LEVEL, "0"), ("flat", N.ACCEL_FLAT, "1"),
                            ("two_serial", N.ACCEL_TWO_LEVEL, "1")):
        _os.environ["SOLB_TLAS_FAST"] = fast  # "0": the multi-kernel TLAS build (instance counts above the single-CTA limit)
        _os.environ["SOLB_TLAS_COOP"] = "0" if key == "two_serial" else "1"  # "0": one thread per wide node (collapse_one)
        try:
            sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=mode)
            for i, t in enumerate(xf):
                sd.add_instance(i % 8, t, i % 8)
            sd.accel_build()
        finally:
            _os.environ.pop("SOLB_TLAS_FAST", None)
            _os.environ.pop("SOLB_TLAS_COOP", None)
        sds[key] = sd
    two, flat = sds["two"], sds["flat"]
    # the warp-cooperative collapse of the single-CTA kernel writes the nodes the one-thread-per-node collapse writes (their
    # numbering depends on the order the warps claim child ranges): same node count and depth, same TLAS bytes once sorted
    ia, ib = two.accel_info(), sds["two_serial"].accel_info()
    assert ia.n_tlas_nodes == ib.n_tlas_nodes and ia.tlas_depth == ib.tlas_depth
    na, nb_ = two.read_nodes()[:ia.n_tlas_nodes].copy(), sds["two_serial"].read_nodes()[:ib.n_tlas_nodes].copy()
    na[:, 4:6] = 0  # child base / leaf base: allocation order
    nb_[:, 4:6] = 0
    assert sorted(map(bytes, na)) == sorted(map(bytes, nb_))
    assert two.accel_info().n_instances == 1000 and two.accel_info().n_blas == 8 and two.accel_info().n_triangles == 32
    assert flat.accel_info().n_triangles == sum(two.instance_triangles())
    n = 300_000
    o = rng.uniform(-25, 25, size=(n, 3))
    d = rng.normal(size=(n, 3))
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)
    h2, t2 = two.trace_rays(rays)
    h1, t1 = flat.trace_rays(rays)
    h3, t3 = sds["two_general"].trace_rays(rays)
    assert np.array_equal(h2, h3) and np.array_equal(t2, t3)  # same BLASes, different TLAS topology: identical hits
    h6, t6 = sds["two_serial"].trace_rays(rays)
    assert np.array_equal(h2, h6) and np.array_equal(t2, t6)
    same = np.all(h2[:, :2] == h1[:, :2], axis=1)
    assert same.mean() > 0.9995  # the two builds round differently only on edge / tie rays
    assert (h1[:, 0] != N.MISS).mean() > 0.2
    dt = np.abs(t2[same] - t1[same]) / (1e-4 + 2e-4 * np.abs(t1[same]))  # grazing hits amplify the object/world rounding difference
    assert (dt > 1).mean() < 1e-4 and dt.max() < 100
    moves = [trs(rng.uniform(-20, 20, 3), (0, 1, 0), 0.1 * k) for k in range(5)]
    for key in ("two", "two_general"):
        times = []
        _os.environ["SOLB_TLAS_FAST"] = "1" if key == "two" else "0"
        try:
            for k in range(5):
                sds[key].blas_transform(moves[k], 100 + k)
                sds[key].tlas_regenerate()
                times.append(ctx.stats().last_build_ms)
        finally:
            _os.environ.pop("SOLB_TLAS_FAST", None)
        print("TLAS regenerate, 1000 instances, %s: %s ms" % ("single-CTA kernel" if key == "two" else "multi-kernel path",
                                                              ", ".join("%.3f" % t for t in times)))
        assert min(times) < 5.0
    h4, _ = two.trace_rays(rays)
    h5, _ = sds["two_general"].trace_rays(rays)
    assert np.array_equal(h4, h5) and not np.array_equal(h4, h2)


@pytest.mark.parametrize("accel", ["two_level", "flat"])
def test_node_graph_instancing_end_to_end(sol, ctx, accel, tmp_path):
    """glTF node-graph instancing: load_scene exposes the further nodes of a mesh, from_scene(instancing=True)
    turns them into instances of the same BLAS; hits and a path-traced frame equal the oracle over the flattened scene."""
    from helpers import write_instanced_gltf
    from oracle import gltf_flatten as gf
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    path = write_instanced_gltf(model_path("Duck"), str(tmp_path / "Duck_inst.gltf"))
    fs = gf.load_scene(path, instancing=True)
    osc = oracle.Scene(fs)
    sc = scene.load_scene(ctx, path)
    sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=N.ACCEL_TWO_LEVEL if accel == "two_level" else N.ACCEL_FLAT, instancing=True)
    info = sd.accel_info()
    assert info.n_instances == 3 and info.n_blas == 1
    assert info.n_triangles == (4212 if accel == "two_level" else 3 * 4212)
    rays = _sphere_rays(osc, 300_000, 21)
    o_hits, o_t, flags = osc.trace_rays(rays, classify=True)
    hits, t = sd.trace_rays(rays)
    assert (np.any(hits[:, :2] != o_hits[:, :2], axis=1) & (flags == 0)).sum() == 0
    assert set(np.unique(o_hits[:, 0]).tolist()) >= {0, 1, 2}
    # one path-traced frame (sky on so the ducks are lit) through the wavefront schedule
    w, h = 160, 120
    cam = scene.Camera((w, h))
    cam.look_at((5, 5, 5), (0, 0, 0), (0, -1, 0))
    ocamera = ocam.Camera((w, h))
    ocamera.look_at((5, 5, 5), (0, 0, 0), (0, -1, 0))
    accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    pathtrace_pipeline(ctx, True).cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, 0), accum, None, max_bounces=8,
                                                                   schedule=N.SCHEDULE_WAVEFRONT), (w, h, 1))
    ref = np.zeros((h, w, 4), np.float32)
    osc.pathtrace_frame(ocam.scene_uniforms(ocamera, w, h, 0), w, h, ref, 0, True, 8, 8)
    d = np.abs(accum.readback()[..., :3] - ref[..., :3])
    assert (d.max(axis=2) > 1e-3 * (1 + ref[..., :3].max(axis=2))).mean() < 0.02 and d.sum() / ref[..., :3].sum() < 0.01


def _tiny_scene(tri_sets, transforms):
    """meshes of hand-made triangles -> (product meshes, oracle FlatScene).  tri_sets[i] = float [k, 3, 3]."""
    from oracle import gltf_flatten as gf
    from sol_rs_b200 import scene

    fs = gf.FlatScene()
    fs.materials = np.array([[0.8, 0.8, 0.8, 1, 0, 0, 0, 0, 0.0, 0.9, 0, 0]], dtype=np.float32)
    meshes, verts, inds, nv, ni = [], [], [], 0, 0
    for tris, t in zip(tri_sets, transforms):
        tris = np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)
        v = np.zeros((tris.shape[0] * 3, 16), dtype=np.float32)
        v[:, 0:3] = tris.reshape(-1, 3)
        v[:, 3] = 1
        v[:, 4:8] = 1
        v[:, 8:12] = [0, 1, 0, 1]
        idx = np.arange(v.shape[0], dtype=np.uint32)
        t = np.asarray(t, dtype=np.float32).reshape(4, 4)
        meshes.append(scene.Mesh("m%d" % len(meshes), v, idx, t.reshape(16), [scene.PrimitiveSection(0, 0, v.shape[0], 0, idx.shape[0], 0)]))
        fs.instances.append(dict(mesh=len(meshes) - 1, transform=t, first_vertex=nv, n_vertices=v.shape[0], first_index=ni,
                                 n_indices=idx.shape[0], material=0))
        verts.append(v)
        inds.append(idx)
        nv += v.shape[0]
        ni += idx.shape[0]
    fs.vertices = np.concatenate(verts)
    fs.indices = np.concatenate(inds)
    return meshes, fs


def test_two_level_edge_cases(sol, ctx):
    """single-triangle BLASes (their own root kernel), a two-triangle BLAS, mirrored / scaled instances, an empty scene:
    two-level == flattened == oracle; a singular instance transform is skipped cleanly."""
    from helpers import trs
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray

    rng = np.random.default_rng(5)
    one = [[[0, 0, 0], [1, 0, 0], [0, 1, 0]]]
    quad = [[[0, 0, 0], [1, 0, 0], [1, 1, 0]], [[0, 0, 0], [1, 1, 0], [0, 1, 0]]]
    blob = rng.uniform(-0.5, 0.5, size=(40, 3, 3))
    cases = {
        "single triangle": ([one], [trs((0, 0, 0))]),
        "two single-triangle BLASes": ([one, one], [trs((0, 0, 0)), trs((0.2, 0.1, 1.0), (0, 0, 1), 0.5)]),
        "mixed": ([one, quad, blob, one], [trs((0, 0, 0)), trs((0, 0, -1), (1, 0, 0), 0.3, (2, 2, 2)), trs((0.3, 0.2, 2.0)),
                                           trs((-0.4, 0, 0.5), (0, 1, 0), 1.0, (-1, 1, 1))]),
    }
    n = 50_000
    o = rng.uniform(-1.5, 1.5, size=(n, 3)) + [0, 0, -4]
    d = np.concatenate([rng.uniform(-0.4, 0.4, size=(n, 2)), np.ones((n, 1))], axis=1)
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)
    for name, (tri_sets, xf) in cases.items():
        meshes, fs = _tiny_scene(tri_sets, xf)
        osc = oracle.Scene(fs)
        o_hits, _, flags = osc.trace_rays(rays, classify=True)
        for mode in (N.ACCEL_TWO_LEVEL, N.ACCEL_FLAT):
            sd = ray.SceneDescription.from_meshes(ctx, meshes, [m.transform for m in meshes], fs.materials, accel_mode=mode)
            hits, _ = sd.trace_rays(rays)
            bad = np.any(hits[:, :2] != o_hits[:, :2], axis=1) & (flags == 0)
            assert bad.sum() == 0, "%s, mode %d: %d unlisted mismatches" % (name, mode, bad.sum())
            info = sd.accel_info()
            assert info.n_instances == len(meshes) and info.mode == mode
        assert (o_hits[:, 0] != oracle.MISS).sum() > 100, name
    # an instance squashed flat by a singular transform cannot be entered in object space (no inverse): the two-level walk
    # skips it instead of producing NaNs; the other instance is unaffected
    meshes, fs = _tiny_scene([blob, quad], [trs((0, 0, 0.5), (0, 0, 1), 0.0, (1, 1, 0)), trs((0, 0, -0.5))])
    sd = ray.SceneDescription.from_meshes(ctx, meshes, [m.transform for m in meshes], fs.materials, accel_mode=N.ACCEL_TWO_LEVEL)
    hits, t = sd.trace_rays(rays)
    assert np.all(hits[:, 0] != 0) and (hits[:, 0] == 1).sum() > 100 and np.all(np.isfinite(t))
    empty = ray.SceneDescription.from_meshes(ctx, [], [], np.zeros((0, 12), np.float32), accel_mode=N.ACCEL_TWO_LEVEL)
    hits, _ = empty.trace_rays(rays[:100])
    assert np.all(hits[:, 0] == N.MISS) and empty.accel_info().n_instances == 0
    empty.tlas_regenerate()


# ---- full-size properties (BASELINE.json sizes; too large for the oracle, checked through invariants) ------------------

@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160)])
def test_full_size_determinism_and_tile_split(sol, ctx, w, h):
    """At the bench sizes: (1) a frame is a pure function of (scene, camera, frame index): two renders are bit-identical
    although queue order and atomics vary; (2) tile split: rendering row tiles of the full-size target one after
    another gives exactly the undivided frame, for ragged tile heights too; (3) ray / path counts agree."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    sc, sd = _product(sol, ctx, "tunnel")
    cam = product_camera(sc, "tunnel", w, h)
    sbt = pathtrace_pipeline(ctx, True)
    u = scene.scene_uniforms(cam, w, h, 3)

    def render(tiles, schedule=N.SCHEDULE_WAVEFRONT):
        accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
        rend = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
        ctx.reset_stats()
        for t in tiles:
            sbt.cmd_trace_rays(ray.TraceBindings(sd, u, accum, rend, accumulation_start_frame=3, max_bounces=8, schedule=schedule, tile_rows=t),
                               (w, h, 1))
        st = ctx.stats()
        return accum.readback(), rend.readback(), int(st.rays), int(st.paths)

    a0, r0, rays0, paths0 = render([None])
    a1, r1, rays1, paths1 = render([None])
    assert np.array_equal(a0, a1) and np.array_equal(r0, r1) and rays0 == rays1
    assert paths0 == w * h * 8 and 5.5 < rays0 / paths0 < 7.0
    third = h // 3 + 1  # ragged: not a multiple of the 4-row pixel tiles
    tiles = [(0, third), (third, third), (2 * third, h - 2 * third)]
    a2, r2, rays2, paths2 = render(tiles)
    assert np.array_equal(a0, a2) and np.array_equal(r0, r2) and rays0 == rays2 and paths0 == paths2
    # interleaved bands (tile_row_stride): 3 "ranks" x 4-row bands, and ragged 6-row bands that straddle the 8x4 pixel tiles
    for band in (4, 6):
        a5, r5, rays5, paths5 = render([(r * band, band, 3 * band) for r in range(3)])
        assert np.array_equal(a0, a5) and np.array_equal(r0, r5) and rays0 == rays5 and paths0 == paths5
    # the warp-local wavefront schedule: deterministic, tile union == undivided frame (contiguous ragged and interleaved bands)
    a7, r7, rays7, paths7 = render([None], N.SCHEDULE_WARPFRONT)
    a8, r8, rays8, paths8 = render(tiles, N.SCHEDULE_WARPFRONT)
    a9, r9, rays9, paths9 = render([(r * 6, 6, 18) for r in (2, 0, 1)], N.SCHEDULE_WARPFRONT)
    assert np.array_equal(a7, a8) and np.array_equal(r7, r8) and rays7 == rays8 and paths7 == paths8 == paths0
    assert np.array_equal(a7, a9) and np.array_equal(r7, r9) and rays7 == rays9
    assert abs(rays7 - rays0) <= 1e-4 * rays0
    if w == 1920:
        a6, r6, rays6, _ = render([(r * 8, 8, 16) for r in (1, 0)], N.SCHEDULE_MEGAKERNEL)
        a3, r3, rays3, _ = render(tiles[::-1], N.SCHEDULE_MEGAKERNEL)
        assert np.array_equal(a3, a6) and rays3 == rays6
        a4, r4, rays4, _ = render([None], N.SCHEDULE_MEGAKERNEL)
        assert np.array_equal(a3, a4) and np.array_equal(r3, r4) and rays3 == rays4
    with pytest.raises(sol.SolbError):
        render([(h - 2, 5)])
    with pytest.raises(sol.SolbError):
        render([(0, 8, 4)])  # bands would overlap


@pytest.mark.parametrize("schedule", [0, 1, 3])
@pytest.mark.parametrize("w,h,spp,mb,frame,start", [(1, 1, 8, 32, 0, 0), (3, 5, 8, 32, 0, 0), (9, 7, 1, 32, 7, 7), (37, 21, 3, 0, 2, 0),
                                                     (64, 33, 8, 1, 0x7FFFFFF0, 0x7FFFFFF0), (130, 70, 16, 4, 1000003, 1000000)])
def test_ragged_sizes_and_extreme_parameters(sol, ctx, w, h, spp, mb, frame, start, schedule):
    """image sizes that do not fill the 8x4 pixel tiles, 1 x 1, one sample, bounce cap 0 (every path shades one hit then
    stops), frame indices near 2^31 (tea seed + alpha = 1 / (frame + 1 - start) arithmetic), a late accumulation start."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    name = "cornell"
    sc, sd = _product(sol, ctx, name)
    cam = product_camera(sc, name, w, h)
    fs, osc = oracle_scene(name)
    ocamera = oracle_camera(fs, name, w, h)
    accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    rend = sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
    sbt = pathtrace_pipeline(ctx, False)
    ref = np.zeros((h, w, 4), np.float32)
    st = oracle.OrcStats()
    ctx.reset_stats()
    for f in (frame, frame + 1):
        sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, rend, accumulation_start_frame=start,
                                             samples_per_frame=spp, max_bounces=mb, schedule=schedule), (w, h, 1))
        o_rgba, _ = osc.pathtrace_frame(ocam.scene_uniforms(ocamera, w, h, f), w, h, ref, start, False, spp, mb, st)
    gs = ctx.stats()
    assert gs.paths == st.paths == 2 * w * h * spp and abs(int(gs.rays) - int(st.rays)) <= max(2, 0.01 * st.rays)
    g = accum.readback()
    d = np.abs(g[..., :3] - ref[..., :3]).max(axis=2)
    bad = d > 1e-3 * (1.0 + np.abs(ref[..., :3]).max(axis=2))
    assert bad.sum() <= max(1, 0.03 * w * h), "%d of %d pixels differ" % (bad.sum(), w * h)
    assert np.all(g[..., 3] == 1.0) and np.all(np.isfinite(g))
    assert (np.abs(rend.readback().astype(np.int32) - o_rgba.astype(np.int32)).max(axis=2) > 1).sum() <= max(1, 0.03 * w * h)


# ---- session-3 knobs: the scheduling variants must not change a single pixel ------------------------------------

def _render_with_env(sol, env, name, w, h, frames, sky, mb, schedule):
    """Render through a private context created under `env` (the SOLB_* knobs are read at context creation)."""
    import os as _os

    saved = {k: _os.environ.get(k) for k in env}
    _os.environ.update(env)
    try:
        c = sol.Context(0)
    finally:
        for k, v in saved.items():
            if v is None:
                _os.environ.pop(k, None)
            else:
                _os.environ[k] = v
    try:
        return _render_gpu(sol, c, name, w, h, frames, sky, 8, mb, schedule)
    finally:
        c.close()


@pytest.mark.parametrize("name,w,h,sky,mb", [("tunnel", 320, 180, True, 8), ("tunnel", 200, 120, True, 32), ("cornell", 128, 128, False, 32)])
def test_async_polls_and_priorities_are_bit_identical(sol, name, w, h, sky, mb):
    """Lock-step polls, asynchronous polls (parts advancing independently), part / shade stream priorities and the run-ahead
    bound only change WHEN the waves of a part are enqueued: every variant must produce the same bits."""
    ref_acc, ref_rgba = _render_with_env(sol, {"SOLB_ASYNC_POLL": "0"}, name, w, h, [0, 1, 2], sky, mb, 0)
    for env in ({"SOLB_ASYNC_POLL": "1"}, {"SOLB_ASYNC_POLL": "1", "SOLB_MAX_AHEAD": "8", "SOLB_CHECK_EVERY": "4"},
                {"SOLB_ASYNC_POLL": "1", "SOLB_PART_PRIORITY": "1"}, {"SOLB_ASYNC_POLL": "1", "SOLB_SHADE_PRIORITY": "1"},
                {"SOLB_ASYNC_POLL": "1", "SOLB_OVERLAP": "4", "SOLB_MIN_PIXELS_PER_PART": "1"}):
        acc, rgba = _render_with_env(sol, env, name, w, h, [0, 1, 2], sky, mb, 0)
        np.testing.assert_array_equal(acc, ref_acc, err_msg=str(env))
        np.testing.assert_array_equal(rgba, ref_rgba, err_msg=str(env))


@pytest.mark.parametrize("name,w,h,sky,mb", [("tunnel", 256, 144, True, 8), ("cornell", 128, 128, False, 32), ("Duck", 160, 120, True, 8)])
def test_voted_traversal_matches_per_thread_traversal(sol, name, w, h, sky, mb):
    """trace_vote (warp-voted steps, postponed triangle groups) against trace_closest (every lane on its own) in the megakernel:
    same closest hits, so the same images except decision-flip pixels (the two instantiations are compiled separately, so FMA
    contraction can differ in the last bit, and the order of equal-t tests differs); same bounds as test_wavefront_equals_megakernel."""
    a, _ = _render_with_env(sol, {"SOLB_MEGA_VOTE": "0"}, name, w, h, [0, 1], sky, mb, 1)
    b, _ = _render_with_env(sol, {"SOLB_MEGA_VOTE": "1"}, name, w, h, [0, 1], sky, mb, 1)
    d = np.abs(a - b)[..., :3]
    assert (d.max(axis=2) > 1e-4 * (1 + np.abs(b[..., :3]).max(axis=2))).mean() < 0.01
    assert d.sum() / b[..., :3].sum() < 2e-3


@pytest.mark.parametrize("env,schedule", [({"SOLB_POOL": "1"}, 0), ({"SOLB_COOP_TRI": "1"}, 0), ({"SOLB_SORT_SHADE": "1"}, 0),
                                          ({"SOLB_MEGA_PERSISTENT": "1"}, 1), ({"SOLB_MEGA_PERSISTENT": "1", "SOLB_MEGA_VOTE": "1"}, 1)])
def test_experimental_schedules_still_agree(sol, env, schedule):
    """The measured-and-switched-off variants (ray-pool kernel, cooperative triangle step, material-sorted shading, persistent
    megakernel) share the node / triangle steps with the default kernels: after any change to those they must still render the
    same image, up to decision-flip pixels (separately compiled instantiations, different order of equal-t tests)."""
    name, w, h = "tunnel", 256, 144
    ref, _ = _render_with_env(sol, {}, name, w, h, [0, 1], True, 8, schedule)
    img, _ = _render_with_env(sol, env, name, w, h, [0, 1], True, 8, schedule)
    d = np.abs(img - ref)[..., :3]
    assert (d.max(axis=2) > 1e-4 * (1 + np.abs(ref[..., :3]).max(axis=2))).mean() < 0.01, str(env)
    assert d.sum() / ref[..., :3].sum() < 2e-3, str(env)


# ---- BASELINE.json sizes against the oracle (configs[0] in full; one full 1080p frame of configs[2]) -------------------------

@pytest.mark.parametrize("schedule", [1, 3])
def test_config0_cornell_512_64spp_cap4_vs_oracle(sol, ctx, schedule):
    """BASELINE.json configs[0] exactly: cornell.gltf 512 x 512, frames 0..7 (64 accumulated spp), bounce cap 4."""
    w = h = 512
    ctx.reset_stats()
    g_acc, g_rgba = _render_gpu(sol, ctx, "cornell", w, h, range(8), False, 8, 4, schedule)
    gs = ctx.stats()
    o_acc, o_rgba, st = _render_oracle("cornell", w, h, range(8), False, 8, 4)
    assert gs.paths == st.paths == w * h * 64
    assert abs(int(gs.rays) - int(st.rays)) <= 0.002 * st.rays
    d = np.abs(g_acc[..., :3] - o_acc[..., :3])
    # 64 samples per pixel: one decision flip moves a pixel by ~1/64 of a sample's colour, so count pixels that moved at all
    assert (d.max(axis=2) > 1e-3 * (1.0 + np.abs(o_acc[..., :3]).max(axis=2))).mean() < 0.05
    mre, psnr = image_metrics(g_acc, o_acc)
    assert mre < 0.002 and psnr > 50.0, (mre, psnr)
    assert (np.abs(g_rgba.astype(np.int32) - o_rgba.astype(np.int32)).max(axis=2) > 1).mean() < 0.01


@pytest.mark.parametrize("schedule", [0, 3])
def test_config2_tunnel_1080p_frame_vs_oracle(sol, ctx, schedule):
    """One full frame of the headline workload (tunnel.gltf --sky 1920 x 1080, 8 spp, bounce cap 8) against the oracle:
    same path count, same ray count within 0.1 %, per-pixel agreement except decision-flip pixels."""
    w, h = 1920, 1080
    ctx.reset_stats()
    g_acc, g_rgba = _render_gpu(sol, ctx, "tunnel", w, h, [0], True, 8, 8, schedule)
    gs = ctx.stats()
    o_acc, o_rgba, st = _render_oracle("tunnel", w, h, [0], True, 8, 8)
    assert gs.paths == st.paths == w * h * 8
    assert abs(int(gs.rays) - int(st.rays)) <= 0.001 * st.rays
    d = np.abs(g_acc[..., :3] - o_acc[..., :3])
    assert (d.max(axis=2) > 1e-3 * (1.0 + np.abs(o_acc[..., :3]).max(axis=2))).mean() < 0.02
    mre, _ = image_metrics(g_acc, o_acc)
    assert mre < 0.01, mre
    assert (np.abs(g_rgba.astype(np.int32) - o_rgba.astype(np.int32)).max(axis=2) > 1).mean() < 0.02


# ---- synthetic instanced scene (BASELINE.json configs[4] at 27 BLAS x 20 000 triangles) -------------------------------------

@pytest.fixture(scope="module")
def synth27():
    from sol_rs_b200 import synth

    sc = synth.make_scene(27, 100)
    fs = flat_from_product_scene(sc)
    return sc, fs, oracle.Scene(fs)


@pytest.mark.parametrize("accel", ["flat", "two_level"])
def test_synth_scene_rays_and_primary_ids_vs_oracle(sol, ctx, synth27, accel):
    """sol_rs_b200.synth (the generator of the 20 M-triangle config) at 27 x 20 000 = 540 000 triangles, the deepest hierarchy
    any test walks: 10^6 incoherent rays and the primary hit ids of its camera, flattened and two-level, against the oracle."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    sc, fs, osc = synth27
    assert osc.tri_count == 27 * 20000
    sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=N.ACCEL_TWO_LEVEL if accel == "two_level" else N.ACCEL_FLAT)
    info = sd.accel_info()
    assert info.n_triangles == osc.tri_count and info.n_instances == 27
    rng = np.random.default_rng(5)
    n = 1_000_000
    lo, hi = osc.bounds()
    o = rng.uniform(lo - 1.0, hi + 1.0, size=(n, 3))
    d = rng.normal(size=(n, 3))
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)
    g_hits, g_t = sd.trace_rays(rays)
    o_hits, o_t, flags = osc.trace_rays(rays, classify=True)
    mism = np.any(g_hits[:, :2] != o_hits[:, :2], axis=1)
    assert (mism & (flags == 0)).sum() == 0, "%d unlisted rays differ" % (mism & (flags == 0)).sum()
    assert (flags != 0).mean() < 0.01
    ok = ~mism & (o_hits[:, 0] != oracle.MISS)
    assert ok.mean() > 0.2
    np.testing.assert_allclose(g_t[ok], o_t[ok], rtol=1e-4, atol=5e-4)
    # primary ids through the debug pipeline with the scene's own camera
    w, h = 960, 540
    cam = sc.camera
    cam.set_window_size((w, h))
    u = scene.scene_uniforms(cam, w, h, 0)
    # a pixel of this view spans about one 0.01-unit triangle seen from 16 units away, often at grazing angles on the bumps:
    # the edge / tie list is taken with 10x the default tolerances (barycentric 1e-3, relative depth gap 1e-4), still < 1 % of pixels
    o_rgba, o_ids, o_bt, o_flags = osc.debug(bytes(u), w, h, 1e-3, 1e-4)
    ids = sol.Image2d(ctx, w, h, N.FORMAT_RG32UI)
    simple_pipeline(ctx, "debug").cmd_trace_rays(ray.TraceBindings(sd, u, None, None, ids), (w, h, 1))
    g_ids = ids.readback()
    mism = np.any(g_ids != o_ids, axis=2)
    assert (mism & (o_flags == 0)).sum() == 0
    assert (o_flags != 0).mean() < 0.01 and (o_ids[..., 0] != oracle.MISS).mean() > 0.3


@pytest.mark.parametrize("schedule", [0, 1, 3])
def test_synth_scene_pathtrace_frame_vs_oracle(sol, ctx, synth27, schedule):
    """Path-traced frames of the synthetic scene (sky on, 8 spp: diffuse, rough-metal and emissive materials).  Its surfaces are
    finely tessellated and curved, so a last-bit difference in a hit position is amplified at every bounce (a dispersing
    billiard): per-pixel agreement with the oracle is only meaningful for short paths (bounce cap 1), full-depth frames are
    compared through their ray statistics and block means."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    sc, fs, osc = synth27
    w, h = 320, 180
    sd = ray.SceneDescription.from_scene(ctx, sc)
    cam = sc.camera
    cam.set_window_size((w, h))
    sbt = pathtrace_pipeline(ctx, True)

    def both(mb, frames):
        accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
        ref = np.zeros((h, w, 4), np.float32)
        st = oracle.OrcStats()
        ctx.reset_stats()
        for f in range(frames):
            u = scene.scene_uniforms(cam, w, h, f)
            sbt.cmd_trace_rays(ray.TraceBindings(sd, u, accum, None, samples_per_frame=8, max_bounces=mb, schedule=schedule), (w, h, 1))
            osc.pathtrace_frame(bytes(u), w, h, ref, 0, True, 8, mb, st)
        return accum.readback(), ref, ctx.stats(), st

    g, ref, gs, st = both(1, 2)  # primary hit + one bounce
    assert gs.paths == st.paths == 2 * 8 * w * h
    assert abs(int(gs.rays) - int(st.rays)) <= 0.002 * st.rays
    d = np.abs(g[..., :3] - ref[..., :3])
    assert (d.max(axis=2) > 1e-3 * (1.0 + np.abs(ref[..., :3]).max(axis=2))).mean() < 0.03
    assert image_metrics(g, ref)[0] < 0.01
    g, ref, gs, st = both(8, 4)  # full depth: statistics
    assert gs.paths == st.paths == 4 * 8 * w * h
    assert abs(int(gs.rays) - int(st.rays)) <= 0.003 * st.rays
    assert abs(int(gs.hits) - int(st.hits)) <= 0.003 * st.hits
    bm = lambda a: a[..., :3].reshape(h // 20, 20, w // 20, 20, 3).mean(axis=(1, 3))  # 20 x 20 pixel block means
    rel = np.abs(bm(g) - bm(ref)) / (bm(ref) + 1e-3)
    assert np.median(rel) < 0.05 and abs(g[..., :3].mean() / ref[..., :3].mean() - 1.0) < 0.02, (np.median(rel), g[..., :3].mean(), ref[..., :3].mean())


# ---- PLOC builder (SOLB_PLOC=1): same invariants and the same hits as the LBVH + treelet build ---------------------------------

@pytest.mark.parametrize("name", ["cornell", "Duck", "tunnel", "synth27"])
def test_ploc_builder_invariants_and_hits(sol, ctx, name, synth27):
    import os as _os2

    from sol_rs_b200 import ray

    if name == "synth27":
        sc, fs, osc = synth27
    else:
        fs, osc = oracle_scene(name)
        from sol_rs_b200 import scene

        sc = scene.load_scene(ctx, model_path(name))
    saved = _os2.environ.get("SOLB_PLOC")
    _os2.environ["SOLB_PLOC"] = "1"
    try:
        sd = ray.SceneDescription.from_scene(ctx, sc)
    finally:
        if saved is None:
            _os2.environ.pop("SOLB_PLOC", None)
        else:
            _os2.environ["SOLB_PLOC"] = saved
    info = sd.accel_info()
    assert info.n_triangles == osc.tri_count and info.sah_cost_binary > 0
    if name != "synth27":  # (the Python walk is too slow for 540 000 triangles)
        seen, n_nodes, depth = _walk_accel(sd.read_nodes(), sd.read_triangles())
        assert np.all(seen == 1) and n_nodes == info.n_wide_nodes and depth == info.wide_depth
    ids = sd.read_triangles().view(np.uint32).reshape(-1, 3, 4)[:, :, 3]
    assert np.array_equal(np.sort(ids[:, 2]), np.arange(osc.tri_count))
    rng = np.random.default_rng(21)
    n = 300_000
    lo, hi = osc.bounds()
    o = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(n, 3))
    d = rng.normal(size=(n, 3))
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)
    g_hits, g_t = sd.trace_rays(rays)
    o_hits, o_t, flags = osc.trace_rays(rays, classify=True)
    mism = np.any(g_hits[:, :2] != o_hits[:, :2], axis=1)
    assert (mism & (flags == 0)).sum() == 0


def test_frames_in_flight_fences_and_async_readback(sol, ctx):
    """A render loop with two frames in flight (src/renderer.rs:72-81, 123-131, 188, 310-317 per-slot fences; the blit of
    examples/5-pathtrace.rs:360-361 as a queued copy): every frame's asynchronous read-back, collected after its fence, holds
    exactly what the blocking read-back of the same frame holds in a loop that waits after every frame."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    w, h, frames = 640, 360, 6
    sc, sd = _product(sol, ctx, "tunnel")
    cam = product_camera(sc, "tunnel", w, h)
    sbt = pathtrace_pipeline(ctx, True)

    def bindings(accum, rend, f):
        return ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, rend, samples_per_frame=4, max_bounces=8,
                                 schedule=N.SCHEDULE_WARPFRONT)

    accum, rend = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F), sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
    blocking = []
    for f in range(frames):
        sbt.cmd_trace_rays(bindings(accum, rend, f), (w, h, 1))
        blocking.append(rend.readback().copy())
    assert any(not np.array_equal(blocking[0], b) for b in blocking[1:])

    accum2, rend2 = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F), sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
    slots = [ctx.host_alloc((h, w, 4), np.uint8) for _ in range(2)]
    fences = [ctx.fence() for _ in range(2)]
    fences[0].wait()  # a new fence is signalled: nothing to wait for
    got = []
    for f in range(frames):
        k = f & 1
        if f >= 2:
            fences[k].wait()
            got.append(slots[k].copy())  # frame f-2, complete once its fence has been waited for
        sbt.cmd_trace_rays(bindings(accum2, rend2, f), (w, h, 1))
        rend2.readback_async(slots[k])
        fences[k].signal()
    for f in range(frames - 2, frames):
        fences[f & 1].wait()
        got.append(slots[f & 1].copy())
    for f in range(frames):
        assert np.array_equal(got[f], blocking[f]), f
    assert np.array_equal(accum.readback(), accum2.readback())


# ---- base-colour textures (an extension shared with the oracle, the reference samples none) ------------------

@pytest.mark.parametrize("schedule", [0, 1, 3])
def test_textured_duck_vs_oracle(sol, ctx, schedule):
    """Duck.gltf with DuckCM.png bound (solb_scene_set_textures): frames agree with the oracle carrying the same extension;
    texture_offset (src/ray/mod.rs:20) reports the material's texture; unbinding restores the untextured frame bit for bit."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    w, h, mb = 160, 120, 8
    sc = scene.load_scene(ctx, model_path("Duck"))
    assert len(sc.textures) == 1 and sc.material_textures == [0] and sc.textures[0].rgba8.shape == (512, 512, 4)
    sd = ray.SceneDescription.from_scene(ctx, sc)
    cam = product_camera(sc, "Duck", w, h)
    sbt = pathtrace_pipeline(ctx, True)

    def render():
        accum, rend = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F), sol.Image2d(ctx, w, h, N.FORMAT_RGBA8)
        for f in range(2):
            sbt.cmd_trace_rays(ray.TraceBindings(sd, scene.scene_uniforms(cam, w, h, f), accum, rend, samples_per_frame=8, max_bounces=mb,
                                                 schedule=schedule), (w, h, 1))
        return accum.readback(), rend.readback()

    plain, _ = render()
    assert all(i.texture_offset == 0 for i in sd.instances())  # nothing bound: the reference's default
    sd.set_textures(sc.textures, sc.material_textures)
    assert [i.texture_offset for i in sd.instances()] == [0] * len(sd.instances())
    g_acc, g_rgba = render()
    fs, osc = oracle_scene("Duck")
    assert np.array_equal(fs.textures[0][0], sc.textures[0].rgba8) and fs.material_textures == [0]
    osc.set_textures(fs.textures, fs.material_textures)
    ocm = oracle_camera(fs, "Duck", w, h)
    o_acc = np.zeros((h, w, 4), dtype=np.float32)
    for f in range(2):
        o_rgba, _ = osc.pathtrace_frame(ocam.scene_uniforms(ocm, w, h, f), w, h, o_acc, 0, True, 8, mb, oracle.OrcStats())
    d = np.abs(g_acc[..., :3] - o_acc[..., :3])
    assert (d.max(axis=2) > 1e-3 * (1.0 + np.abs(o_acc[..., :3]).max(axis=2))).mean() < 0.02
    assert image_metrics(g_acc, o_acc)[0] < 0.01
    assert (np.abs(g_rgba.astype(np.int32) - o_rgba.astype(np.int32)).max(axis=2) > 1).mean() < 0.02
    # the texture did something (the duck is yellow, not white), on the duck's pixels only
    changed = np.abs(g_acc[..., :3] - plain[..., :3]).max(axis=2) > 1e-3
    assert 0.05 < changed.mean() < 0.9
    assert g_acc[changed][:, 2].mean() < 0.7 * plain[changed][:, 2].mean()
    sd.set_textures([], [None] * len(sc.material_textures))
    again, _ = render()
    assert np.array_equal(again, plain)


@pytest.mark.parametrize("wrap", [10497, 33071, 33648])
def test_texture_sampling_and_wrap_modes_vs_oracle(sol, ctx, wrap):
    """A quad whose uv run from -1.3 to 2.4 under a 5 x 3 random texture: bilinear taps, texel-centre convention, the three glTF
    wrap modes and the sRGB decode, against the oracle's restatement (primary hit x sky: one bounce)."""
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    w, h = 192, 128
    rng = np.random.default_rng(wrap)
    tex = rng.integers(0, 256, size=(3, 5, 4), dtype=np.uint8)
    v = np.zeros((4, 16), dtype=np.float32)
    for k, (x, y) in enumerate([(-1, -1), (1, -1), (1, 1), (-1, 1)]):
        v[k, 0:4] = (x, y, 0, 1)
        v[k, 4:8] = (1, 1, 1, 1)
        v[k, 8:12] = (0, 0, 1, 1)
        v[k, 12:14] = (-1.3 + 3.7 * (x + 1) / 2, -0.8 + 2.9 * (y + 1) / 2)
    mesh = scene.Mesh("quad", v, np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32), np.eye(4, dtype=np.float32).reshape(16),
                      [scene.PrimitiveSection(0, 0, 4, 0, 6, 0)])
    mats = np.array([[1, 1, 1, 1, 0, 0, 0, 0, 0.0, 1.0, 0, 0]], dtype=np.float32)
    sc = scene.Scene([mesh], mats, None, [scene.Texture(tex, wrap, wrap)], [0])
    sd = ray.SceneDescription.from_scene(ctx, sc, textures=True)
    cam = scene.Camera((w, h))
    cam.look_at((0.3, 0.2, 3.0), (0, 0, 0), (0, -1, 0))
    accum = sol.Image2d(ctx, w, h, N.FORMAT_RGBA32F)
    u = scene.scene_uniforms(cam, w, h, 0)
    pathtrace_pipeline(ctx, True).cmd_trace_rays(ray.TraceBindings(sd, u, accum, None, samples_per_frame=8, max_bounces=1), (w, h, 1))
    g = accum.readback()
    fs = flat_from_product_scene(sc)
    osc = oracle.Scene(fs)
    osc.set_textures([(tex, wrap, wrap)], [0])
    ref = np.zeros((h, w, 4), dtype=np.float32)
    osc.pathtrace_frame(bytes(u), w, h, ref, 0, True, 8, 1, oracle.OrcStats())
    d = np.abs(g[..., :3] - ref[..., :3])
    assert (d.max(axis=2) > 1e-4 * (1.0 + np.abs(ref[..., :3]).max(axis=2))).mean() < 0.005
    assert image_metrics(g, ref)[0] < 1e-3
    # and the oracle's sampler is what the header says: texel centres at (i + 0.5) / n, sRGB decode, wrap
    lin = lambda c: np.where(c / 255.0 <= 0.04045, c / 255.0 / 12.92, ((c / 255.0 + 0.055) / 1.055) ** 2.4)
    np.testing.assert_allclose(osc.sample_texture(0, 1.5 / 5, 2.5 / 3), lin(tex[2, 1, :3].astype(np.float64)), rtol=1e-5, atol=1e-6)
    edge = osc.sample_texture(0, 1.0, 0.5 / 3)  # u = 1: between the last and (repeat) the first / (clamp, mirror) the last texel
    want = 0.5 * (lin(tex[0, 4, :3].astype(np.float64)) + lin(tex[0, 0 if wrap == 10497 else 4, :3].astype(np.float64)))
    np.testing.assert_allclose(edge, want, rtol=1e-5, atol=1e-6)
