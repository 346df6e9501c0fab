#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/r2_hi_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_HI_STREAM=1
run SOLB_HI_STREAM=0
run SOLB_HI_STREAM=1
for v in "SOLB_HI_STREAM=1" "SOLB_HI_STREAM=0"; do for w in 8; do echo -n "$v : "; env $v timeout 120 python tools/tile_time.py $w 2>&1 | head -1; done; done
} > gpurun_out/r2_hi_bench.log 2>&1
tail -6 gpurun_out/r2_hi_tests.log; cat gpurun_out/r2_hi_bench.log
