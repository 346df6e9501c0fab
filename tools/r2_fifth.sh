#!/bin/bash
# 2-GPU box: comm tests, scaling bench line, 1-GPU A/B of the FIFO lists
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_comm.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_fifth_tests.log 2>&1
{
run() { echo -n "$* -> "; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['schedule'], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; }
run SOLB_SCHEDULE=wavefront
run SOLB_SCHEDULE=warpfront
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=8
run SOLB_SCHEDULE=warpfront SOLB_WL_FETCH_IDLE=16
} > gpurun_out/r2_fifth_bench.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_fifth_2gpu.json 2> gpurun_out/r2_fifth_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 8 --warmup 3 --schedule wavefront > gpurun_out/r2_fifth_2gpu_wf.json 2>> gpurun_out/r2_fifth_2gpu.err
tail -5 gpurun_out/r2_fifth_tests.log; cat gpurun_out/r2_fifth_bench.log; cut -c1-600 gpurun_out/r2_fifth_2gpu.json; python -c "
import json
for f in ('gpurun_out/r2_fifth_2gpu.json','gpurun_out/r2_fifth_2gpu_wf.json'):
    d=json.loads(open(f).read()); print(d['config']['schedule'], d['value'], d['e2e']['value'], d['reduce_ms'], d['extra'])
"; tail -5 gpurun_out/r2_fifth_2gpu.err
