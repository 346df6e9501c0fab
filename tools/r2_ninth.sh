#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_ninth_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'nodes/tris per ray', round(d['roofline']['nodes_per_ray'],2), round(d['roofline']['tris_per_ray'],2), 'build', round(d['config']['bvh_build_ms'],2), round(d['config']['bvh_rebuild_ms'],2))"; }
run SOLB_SCHEDULE=warpfront
run SOLB_SCHEDULE=warpfront SOLB_PLOC=1 SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_plocr16.so
echo "== TLAS regen"
SOLB_TLAS_TRACE=1 python tools/tlas_regen_bench.py --frames 6 2>&1 | tail -4
SOLB_TLAS_COOP=0 python tools/tlas_regen_bench.py --frames 6 2>&1 | tail -1
echo "== synth 20M: treelets / PLOC r8 / PLOC r16 (bench --workload synth: build ms first + warm rebuild)"
for v in "SOLB_X=0" "SOLB_PLOC=1" "SOLB_PLOC=1 SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_plocr16.so"; do
  echo -n "$v -> "; env $v timeout 600 python bench.py --workload synth --steps 4 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', 'nodes/tris', round(d['roofline']['nodes_per_ray'],2), round(d['roofline']['tris_per_ray'],2), 'build first/warm', round(d['config']['bvh_build_ms'],1), round(d['config']['bvh_rebuild_ms'],1), 'frac', round(d['roofline']['frac'],3))"
done
} > gpurun_out/r2_ninth_bench.log 2>&1
tail -6 gpurun_out/r2_ninth_tests.log; cat gpurun_out/r2_ninth_bench.log
