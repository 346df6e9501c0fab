#!/bin/bash
mkdir -p gpurun_out
{
for v in "SOLB_X=0" "SOLB_WL_FETCH_IDLE_LARGE=4" "SOLB_WL_FETCH_IDLE_LARGE=6" "SOLB_WL_FETCH_IDLE_LARGE=8 SOLB_WL_STARVE_IDLE=8" "SOLB_WL_FETCH_IDLE_LARGE=10" "SOLB_WL_FETCH_IDLE_LARGE=20"; do
  echo -n "$v -> "; env $v timeout 300 python bench.py --workload synth --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'launch ms', round(d['roofline']['avg_launch_ms'],2))"
done
echo -n "tunnel default -> "; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3))"
} > gpurun_out/r2_synth_knobs2.log 2>&1
cat gpurun_out/r2_synth_knobs2.log
