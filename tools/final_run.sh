#!/bin/bash
# round-end evidence run on ONE GPU (gpurun -- bash tools/final_run.sh; the multi-GPU lines: torchrun bench.py --gpus N): full GPU test suite, smoke, the default bench line (all extras + CPU baseline), the reference
# arm, an ncu launch list and one full capture of the dominant kernel
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r02b_gpu_tests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > gpurun_out/r02b_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/r02b_bench_reference.json 2>> gpurun_out/r02b_bench_default.err
timeout 600 python bench.py --steps 8 --warmup 3 --schedule wavefront --no-cpu-baseline --no-extra > gpurun_out/r02b_bench_wavefront.json 2>> gpurun_out/r02b_bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02b_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pt_warpfront -s 4 -c 1 -f -o gpurun_out/r02b_prof_warpfront python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02b_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pt_warpfront -s 3 -c 1 -f -o gpurun_out/r02b_prof_warpfront_synth python bench.py --workload synth --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02b_ncu_full_synth.log 2>&1
cat gpurun_out/r02b_gpu_tests.log gpurun_out/r02b_smoke.log; cut -c1-400 gpurun_out/r02b_bench_default.json; tail -3 gpurun_out/r02b_bench_default.err
