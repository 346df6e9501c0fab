"""debug aid: GPU vs oracle on the synthetic scene (which rays differ, and is every triangle reachable)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from helpers import flat_from_product_scene
import sol_rs_b200 as sol
from sol_rs_b200 import synth, ray, _native as N

n_blas = int(sys.argv[1]) if len(sys.argv) > 1 else 27
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ctx = sol.Context(0)
sc = synth.make_scene(n_blas, grid)
fs = flat_from_product_scene(sc)
osc = oracle.Scene(fs)
for mode in (N.ACCEL_FLAT, N.ACCEL_TWO_LEVEL):
    sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=mode)
    info = sd.accel_info()
    print("mode", mode, "tris", info.n_triangles, "wide", info.n_wide_nodes, "depth", info.wide_depth, "sah", info.sah_cost_binary, info.sah_cost_lbvh)
    tris = sd.read_triangles()
    ids = tris.view(np.uint32).reshape(-1, 3, 4)[:, :, 3]
    g = np.sort(ids[:, 2])
    print("  ordinals unique:", np.array_equal(g, np.arange(len(g))), "n", len(g))
    rng = np.random.default_rng(5)
    n = 200000
    lo, hi = osc.bounds()
    o = rng.uniform(lo - 1, hi + 1, size=(n, 3)); d = rng.normal(size=(n, 3))
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], axis=1).astype(np.float32)
    gh, gt = sd.trace_rays(rays)
    oh, ot, fl = osc.trace_rays(rays, classify=True)
    mism = np.any(gh[:, :2] != oh[:, :2], axis=1)
    print("  mismatch", mism.sum(), "unlisted", (mism & (fl == 0)).sum(), "gpu miss/oracle hit", ((gh[:, 0] == oracle.MISS) & (oh[:, 0] != oracle.MISS)).sum(),
          "gpu hit/oracle miss", ((gh[:, 0] != oracle.MISS) & (oh[:, 0] == oracle.MISS)).sum(), "both hit differ", (mism & (gh[:, 0] != oracle.MISS) & (oh[:, 0] != oracle.MISS)).sum())
    idx = np.where(mism & (fl == 0))[0][:8]
    for i in idx:
        print("   ", i, gh[i, :2], oh[i, :2], gt[i], ot[i])
