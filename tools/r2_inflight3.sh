#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "frames_in_flight or warpfront_equals or determin" 2>&1 | tail -8 ) > gpurun_out/r2_inflight3_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_X=0
run SOLB_WL_FRAMES_IN_FLIGHT=3
run SOLB_WL_FRAMES_IN_FLIGHT=3 SOLB_WL_WARPS_PER_SM=30
run SOLB_WL_FRAMES_IN_FLIGHT=2 SOLB_WL_WARPS_PER_SM=30
run SOLB_WL_FRAMES_IN_FLIGHT=2 SOLB_WL_WARPS_PER_SM=28
run SOLB_WL_FRAMES_IN_FLIGHT=3 SOLB_WL_FETCH_IDLE=12
run SOLB_WL_FRAMES_IN_FLIGHT=3 SOLB_WL_FETCH_IDLE=20 SOLB_WL_STARVE_IDLE=20
run SOLB_WL_FRAMES_IN_FLIGHT=3
echo "== synth 20M"
for v in "SOLB_WL_FRAMES_IN_FLIGHT=1" "SOLB_WL_FRAMES_IN_FLIGHT=2" "SOLB_WL_FRAMES_IN_FLIGHT=3"; do
  echo -n "$v -> "; env $v timeout 600 python bench.py --workload synth --steps 6 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', 'e2e', round(d['e2e']['value'],1), 'build first/warm', round(d['config']['bvh_build_ms'],1), round(d['config']['bvh_rebuild_ms'],1), 'frac', round(d['roofline']['frac'],3))"
done
} > gpurun_out/r2_inflight3_bench.log 2>&1
tail -8 gpurun_out/r2_inflight3_tests.log; cat gpurun_out/r2_inflight3_bench.log
