"""Phase-by-phase wall clock of the acceleration-structure build on the synthetic instanced scene (SOLB_BUILD_TRACE).
    python tools/build_trace.py [n_blas]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SOLB_BUILD_TRACE"] = "1"
import sol_rs_b200 as sol  # noqa: E402
from sol_rs_b200 import _native as N, ray, synth  # noqa: E402

ctx = sol.Context(0)
sc = synth.make_scene(int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 100)
for dp in ("1", "0"):
    os.environ["SOLB_DP_COLLAPSE"] = dp
    for mode, name in ((N.ACCEL_FLAT, "flat"), (N.ACCEL_TWO_LEVEL, "two_level"), (N.ACCEL_FLAT, "flat again")):
        t0 = time.perf_counter()
        sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=mode)
        info = sd.accel_info()
        print("== %s dp=%s: from_scene %.1f ms, last_build_ms %.1f, wide nodes %d, depth %d" % (
            name, dp, 1e3 * (time.perf_counter() - t0), ctx.stats().last_build_ms, info.n_wide_nodes, info.wide_depth), file=sys.stderr)
        sd.close()
