#!/bin/bash
mkdir -p gpurun_out
{
for lib in "" "stk6" "stk4"; do
  L=""; [ -n "$lib" ] && L="SOLB_LIB_PATH=$PWD/sol_rs_b200/libsolb_$lib.so"
  echo -n "synth ${lib:-default(8)} -> "; env $L SOLB_X=0 timeout 300 python bench.py --workload synth --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', 'frac', round(d['roofline']['frac'],3), 'launch ms', round(d['roofline']['avg_launch_ms'],2))"
  echo -n "tunnel ${lib:-default(8)} -> "; env $L SOLB_X=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3))"
done
} > gpurun_out/r2_stack.log 2>&1
cat gpurun_out/r2_stack.log
