#!/bin/bash
mkdir -p gpurun_out
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_X=0
run SOLB_WL_FETCH_IDLE=12
run SOLB_WL_FETCH_IDLE=8
run SOLB_WL_FETCH_IDLE=4
run SOLB_WL_FETCH_IDLE=8 SOLB_WL_STARVE_IDLE=24
run SOLB_WL_FETCH_IDLE=16 SOLB_WL_STARVE_IDLE=24
run SOLB_WL_FETCH_IDLE=16 SOLB_WL_STARVE_IDLE=32
} > gpurun_out/r2_quick_bench.log 2>&1
cat gpurun_out/r2_quick_bench.log
