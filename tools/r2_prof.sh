#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pt_warpfront -s 4 -c 1 -f -o gpurun_out/prof_warpfront_d python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_warpfront_d.log 2>&1
ls -la gpurun_out/prof_warpfront_d.ncu-rep
