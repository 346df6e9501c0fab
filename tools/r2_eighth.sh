#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_eighth_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'nodes/tris per ray', round(d['roofline']['nodes_per_ray'],2), round(d['roofline']['tris_per_ray'],2), 'build', round(d['config']['bvh_build_ms'],2))"; }
run SOLB_SCHEDULE=wavefront
run SOLB_SCHEDULE=warpfront
run SOLB_SCHEDULE=warpfront SOLB_PLOC=1
run SOLB_SCHEDULE=wavefront SOLB_PLOC=1
echo "== build trace 20M, LBVH+treelets vs PLOC"
SOLB_BUILD_TRACE=1 timeout 300 python tools/build_trace.py 1000 2>&1 | grep -v "two_level" | head -40
echo "== PLOC"
SOLB_PLOC=1 SOLB_BUILD_TRACE=1 timeout 300 python tools/build_trace.py 1000 2>&1 | head -40
} > gpurun_out/r2_eighth_bench.log 2>&1
tail -8 gpurun_out/r2_eighth_tests.log; cat gpurun_out/r2_eighth_bench.log
