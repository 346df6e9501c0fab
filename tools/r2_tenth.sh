#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_tenth_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_SCHEDULE=warpfront
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=12
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=8
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=4
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=8 SOLB_WL_GEN_MIN=16
echo "== TLAS regen"
SOLB_TLAS_TRACE=1 python tools/tlas_regen_bench.py --frames 6 2>&1 | tail -4
} > gpurun_out/r2_tenth_bench.log 2>&1
tail -6 gpurun_out/r2_tenth_tests.log; cat gpurun_out/r2_tenth_bench.log
