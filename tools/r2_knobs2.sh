#!/bin/bash
mkdir -p gpurun_out
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_X=0
run SOLB_WL_FETCH_IDLE=20 SOLB_WL_STARVE_IDLE=20
run SOLB_WL_FETCH_IDLE=24 SOLB_WL_STARVE_IDLE=24
run SOLB_WL_FETCH_IDLE=20 SOLB_WL_STARVE_IDLE=16
run SOLB_WL_FETCH_IDLE=16 SOLB_WL_STARVE_IDLE=24
run SOLB_WL_GEN_MIN=24
run SOLB_WL_GEN_MIN=48
run SOLB_WL_BATCH=64
run SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_p128.so
run SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_p128.so SOLB_WL_FETCH_IDLE=20 SOLB_WL_STARVE_IDLE=20
run SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_p64.so
run SOLB_WL_FRAMES_IN_FLIGHT=2
run SOLB_X=0
} > gpurun_out/r2_knobs2.log 2>&1
cat gpurun_out/r2_knobs2.log
