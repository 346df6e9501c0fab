#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2_fourth_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_SCHEDULE=wavefront
run SOLB_SCHEDULE=warpfront
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=12 SOLB_WL_GEN_MIN=32
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=16 SOLB_WL_GEN_MIN=32
run SOLB_SCHEDULE=warpfront SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_r72.so SOLB_WL_CTAS_PER_SM=7
} > gpurun_out/r2_fourth_bench.log 2>&1
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_fourth_full.json 2> gpurun_out/r2_fourth_full.err
tail -6 gpurun_out/r2_fourth_tests.log; cat gpurun_out/r2_fourth_bench.log; cut -c1-1500 gpurun_out/r2_fourth_full.json; tail -3 gpurun_out/r2_fourth_full.err
