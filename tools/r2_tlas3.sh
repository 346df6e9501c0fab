#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "two_level or tlas or instanc or node_graph" 2>&1 | tail -6 ) > gpurun_out/r2_tlas3_tests.log 2>&1
{
SOLB_TLAS_TRACE=1 timeout 120 python tools/tlas_regen_bench.py --frames 6 2>&1 | tail -4 | cut -c1-330
for n in 64 256 2000; do timeout 120 python tools/tlas_regen_bench.py --frames 6 --instances $n 2>&1 | tail -1 | cut -c1-200; done
} > gpurun_out/r2_tlas3.log 2>&1
tail -6 gpurun_out/r2_tlas3_tests.log; cat gpurun_out/r2_tlas3.log
