#!/bin/bash
# quick scheduling sweep of the wavefront trace kernel (tunnel 1080p cap 8)
run() { echo -n "$* -> "; env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms', 'trace_launch_ms', round(d['roofline']['avg_launch_ms'],4))"; }
run SOLB_FETCH_IDLE=8
run SOLB_FETCH_IDLE=4
run SOLB_FETCH_IDLE=2
run SOLB_FETCH_IDLE=12
run SOLB_FETCH_IDLE=16
run SOLB_FETCH_IDLE=8 SOLB_TRI_WEIGHT=2
run SOLB_FETCH_IDLE=8 SOLB_TRI_WEIGHT=3
run SOLB_FETCH_IDLE=8 SOLB_NODE_WEIGHT=2
run SOLB_FETCH_IDLE=8 SOLB_CTAS_PER_SM=4
run SOLB_FETCH_IDLE=8 SOLB_CTAS_PER_SM=6
run SOLB_FETCH_IDLE=8 SOLB_CTAS_PER_SM=7
run SOLB_FETCH_IDLE=8 SOLB_CHECK_EVERY=16
run SOLB_FETCH_IDLE=8 SOLB_CHECK_EVERY=4
