#!/bin/bash
# frames in flight A/B: tests first (the schedule-3 ones and everything that renders through AUTO), then bench both ways
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 ) > gpurun_out/r2_inflight_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_WL_FRAMES_IN_FLIGHT=1
run SOLB_WL_FRAMES_IN_FLIGHT=2
run SOLB_WL_FRAMES_IN_FLIGHT=2 SOLB_WL_CTAS_PER_SM=7
run SOLB_WL_FRAMES_IN_FLIGHT=2 SOLB_WL_CTAS_PER_SM=6
run SOLB_WL_FRAMES_IN_FLIGHT=1
run SOLB_WL_FRAMES_IN_FLIGHT=2
} > gpurun_out/r2_inflight_bench.log 2>&1
tail -8 gpurun_out/r2_inflight_tests.log; cat gpurun_out/r2_inflight_bench.log
