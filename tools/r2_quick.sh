#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warpfront or ragged or two_level_pathtrace or frames_vs_oracle or experimental or wavefront_equals or voted or async" 2>&1 | tail -6 ) > gpurun_out/r2_quick_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_X=0
run SOLB_SCHEDULE=wavefront
run SOLB_SCHEDULE=megakernel
run SOLB_X=1
} > gpurun_out/r2_quick_bench.log 2>&1
tail -4 gpurun_out/r2_quick_tests.log; cat gpurun_out/r2_quick_bench.log
