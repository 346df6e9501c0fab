"""TLAS::regenerate cost in two-level mode: N instances of cornell's 8 BLASes, one instance moved per
frame, `solb_tlas_regenerate` timed (device events inside the library + host wall clock around the call).
    python tools/tlas_regen_bench.py [--instances 1000] [--frames 50]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=1000)
    ap.add_argument("--frames", type=int, default=50)
    args = ap.parse_args()
    import sol_rs_b200 as sol
    from helpers import trs
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import ray, scene

    ctx = sol.Context(0)
    sc = scene.load_scene(ctx, os.path.join(ROOT, "assets", "models", "cornell.gltf"))
    rng = np.random.default_rng(3)
    out = {"instances": args.instances, "frames": args.frames}
    for label, fast in (("single_cta", "1"), ("multi_kernel", "0")):
        os.environ["SOLB_TLAS_FAST"] = fast
        sd = ray.SceneDescription.from_scene(ctx, sc, accel_mode=N.ACCEL_TWO_LEVEL)
        for i in range(args.instances - 8):
            sd.add_instance(i % 8, trs(rng.uniform(-20, 20, 3), rng.normal(size=3), rng.uniform(0, 6.28), (rng.uniform(0.5, 1.5),) * 3), i % 8)
        sd.accel_build()
        dev, wall = [], []
        for k in range(args.frames):
            sd.blas_transform(trs(rng.uniform(-20, 20, 3), (0, 1, 0), 0.1 * k), k % args.instances)
            t0 = time.perf_counter()
            sd.tlas_regenerate()
            wall.append(1e3 * (time.perf_counter() - t0))
            dev.append(ctx.stats().last_build_ms)
        out[label] = {"device_ms_median": float(np.median(dev)), "device_ms_min": float(min(dev)),
                      "host_call_ms_median": float(np.median(wall)), "tlas_nodes": sd.accel_info().n_tlas_nodes,
                      "tlas_depth": sd.accel_info().tlas_depth}
    os.environ.pop("SOLB_TLAS_FAST", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
