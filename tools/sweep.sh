#!/bin/bash
# quick sweep of wavefront knobs (tunnel 1080p cap 8): bash tools/sweep.sh "VAR=a VAR=b ..." each arg one run
run() { echo -n "$* -> "; env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms', 'trace_launch_ms', round(d['roofline']['avg_launch_ms'],4), 'launches', d['gpu_launches'])"; }
for cfg in "$@"; do run $cfg; done
