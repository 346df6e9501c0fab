#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_parity.py -x -q -m gpu -k "comm or two_ranks or world_one or warpfront or full_size or ragged" 2>&1 | tail -12 ) > gpurun_out/r2_sixth_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_SCHEDULE=warpfront
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=20
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=24
SOLB_TLAS_TRACE=1 python tools/tlas_regen_bench.py --frames 6 2>&1 | tail -8
} > gpurun_out/r2_sixth_bench.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_sixth_2gpu.json 2> gpurun_out/r2_sixth_2gpu.err
tail -5 gpurun_out/r2_sixth_tests.log; cat gpurun_out/r2_sixth_bench.log; python -c "
import json
for f in ('gpurun_out/r2_sixth_2gpu.json',):
    d=json.loads(open(f).read()); print(d['config']['schedule'], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'reduce_ms', d['reduce_ms'], d['extra'], d['gpu_launches'])
"; tail -3 gpurun_out/r2_sixth_2gpu.err
