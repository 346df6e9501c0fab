#!/bin/bash
mkdir -p gpurun_out
{
for v in "SOLB_X=0" "SOLB_WL_FETCH_IDLE=12 SOLB_WL_STARVE_IDLE=12" "SOLB_WL_FETCH_IDLE=16" "SOLB_WL_FETCH_IDLE=24 SOLB_WL_STARVE_IDLE=24" "SOLB_WL_FETCH_IDLE=8 SOLB_WL_STARVE_IDLE=16" "SOLB_WL_GEN_MIN=48" "SOLB_WL_GEN_MIN=16"; do
  echo -n "$v -> "; env $v timeout 300 python bench.py --workload synth --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'launch ms', round(d['roofline']['avg_launch_ms'],2))"
done
} > gpurun_out/r2_synth_knobs.log 2>&1
cat gpurun_out/r2_synth_knobs.log
