#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warpfront or ragged" 2>&1 | tail -5 ) > gpurun_out/r2_second_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --schedule warpfront 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1))"; }
run SOLB_X=0
run SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_cg0.so
run SOLB_WL_FETCH_IDLE=12
run SOLB_WL_FETCH_IDLE=16
run SOLB_WL_GEN_MIN=32
run SOLB_WL_GEN_MIN=32 SOLB_WL_FETCH_IDLE=12
run SOLB_TRI_WEIGHT=2
} > gpurun_out/r2_second_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pt_warpfront -s 4 -c 1 -f -o gpurun_out/prof_warpfront_a python bench.py --steps 2 --warmup 3 --no-cpu-baseline --schedule warpfront > gpurun_out/ncu_warpfront_a.log 2>&1
tail -3 gpurun_out/r2_second_tests.log; cat gpurun_out/r2_second_bench.log
