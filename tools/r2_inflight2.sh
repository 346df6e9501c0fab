#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "frames_in_flight or warpfront_equals or determin" 2>&1 | tail -8 ) > gpurun_out/r2_inflight2_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_X=0
run SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_wlb64.so
run SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_wlb32.so
run SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_wlb32.so SOLB_WL_FRAMES_IN_FLIGHT=1
run SOLB_X=0
} > gpurun_out/r2_inflight2_bench.log 2>&1
tail -8 gpurun_out/r2_inflight2_tests.log; cat gpurun_out/r2_inflight2_bench.log
