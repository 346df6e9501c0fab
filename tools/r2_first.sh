#!/bin/bash
# round-2 first GPU pass: warp-local wavefront correctness + first numbers
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warpfront or (frames_vs_oracle and 3) or ragged" 2>&1 | tail -15 ) > gpurun_out/r2_first_tests.log 2>&1
{
for s in wavefront warpfront; do
  echo "== schedule $s"; timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --schedule $s 2>/dev/null
done
for v in "SOLB_WL_FETCH_IDLE=4" "SOLB_WL_FETCH_IDLE=12" "SOLB_WL_GEN_MIN=8" "SOLB_WL_GEN_MIN=32" "SOLB_WL_CTAS_PER_SM=6" "SOLB_WL_CTAS_PER_SM=7" "SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_p64.so" "SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_p128.so"; do
  echo "== $v"; env $v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --schedule warpfront 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1))"
done
} > gpurun_out/r2_first_bench.log 2>&1
tail -5 gpurun_out/r2_first_tests.log; cat gpurun_out/r2_first_bench.log | cut -c1-400
