import os, sys, time
sys.path.insert(0, '/root/repo')
os.environ["SOLB_BUILD_TRACE"] = "1"
import sol_rs_b200 as sol
from sol_rs_b200 import _native as N, ray, synth
ctx = sol.Context(0)
sc = synth.make_scene(int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 100)
for mode, name in ((N.ACCEL_FLAT, "flat"), (N.ACCEL_TWO_LEVEL, "two_level"), (N.ACCEL_FLAT, "flat again"), (N.ACCEL_TWO_LEVEL, "two_level again")):
    t0 = time.perf_counter()
    sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=mode)
    print("== %s: from_scene %.1f ms, last_build_ms %.1f" % (name, 1e3 * (time.perf_counter() - t0), ctx.stats().last_build_ms), file=sys.stderr)
    sd.close()
