#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "build or ploc or rebuild or accel or two_level or synth or update or instance" 2>&1 | tail -6 ) > gpurun_out/r2_build2_tests.log 2>&1
{
for v in "SOLB_X=0" "SOLB_X=1"; do
  echo -n "$v -> "; env $v timeout 600 python bench.py --workload synth --steps 6 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', 'e2e', round(d['e2e']['value'],1), 'build first/warm', round(d['config']['bvh_build_ms'],1), round(d['config']['bvh_rebuild_ms'],1), 'frac', round(d['roofline']['frac'],3))"
done
SOLB_BUILD_TRACE=1 timeout 300 python tools/build_trace.py 1000 2>&1 | tail -30
} > gpurun_out/r2_build2.log 2>&1
tail -6 gpurun_out/r2_build2_tests.log; cat gpurun_out/r2_build2.log
