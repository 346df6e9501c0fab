#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warpfront_equals or full_size or synth" 2>&1 | tail -12 ) > gpurun_out/r2_seventh_tests.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "warpfront_equals and 1920" --count 1 2>&1 | tail -4 ) >> gpurun_out/r2_seventh_tests.log 2>&1
for i in 1 2 3; do ( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "warpfront_equals and 1920" 2>&1 | tail -2 ) >> gpurun_out/r2_seventh_tests.log 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pt_warpfront -s 4 -c 1 -f -o gpurun_out/prof_warpfront_c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --schedule warpfront > gpurun_out/ncu_warpfront_c.log 2>&1
cat gpurun_out/r2_seventh_tests.log
