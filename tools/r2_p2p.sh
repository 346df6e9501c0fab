#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
( timeout 300 python -m pytest tests/test_gpu_comm.py -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/r2_p2p_tests.log 2>&1
tail -15 gpurun_out/r2_p2p_tests.log
bash tools/r2_tile8.sh $N SOLB_P2P=1 SOLB_P2P=0
