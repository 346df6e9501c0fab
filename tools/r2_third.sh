#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warpfront or ragged or config0 or config2 or synth or accel_invariants" 2>&1 | tail -8 ) > gpurun_out/r2_third_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_SCHEDULE=wavefront
run SOLB_SCHEDULE=warpfront
run SOLB_SCHEDULE=warpfront SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_stk4.so
run SOLB_SCHEDULE=warpfront SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_cg0.so
run SOLB_SCHEDULE=warpfront SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_p64.so
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=12
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=4
run SOLB_SCHEDULE=warpfront SOLB_WL_GEN_MIN=32
run SOLB_SCHEDULE=wavefront
} > gpurun_out/r2_third_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pt_warpfront -s 4 -c 1 -f -o gpurun_out/prof_warpfront_b python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --schedule warpfront > gpurun_out/ncu_warpfront_b.log 2>&1
tail -4 gpurun_out/r2_third_tests.log; cat gpurun_out/r2_third_bench.log
