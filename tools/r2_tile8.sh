#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
shift
{
for v in "$@"; do
  echo "== $v"
  env $v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/tile_split_bench.py 48 2>gpurun_out/r2_tile8.err | grep world
done
} > gpurun_out/r2_tile8_$N.log 2>&1
cat gpurun_out/r2_tile8_$N.log
