run() { echo -n "$* -> "; env "$@" python bench.py --workload synth --blas 1000 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms')"; }
for cfg in "$@"; do run $cfg; done
