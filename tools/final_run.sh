#!/bin/bash
# round-end evidence run on a GPU box: full GPU test suite, bench lines, ncu launch list + one full capture of the dominant kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/gpu_tests_final.log 2>&1
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final.log 2> gpurun_out/bench_reference_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 30 -c 1 -f -o gpurun_out/prof_wf_trace_r01_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
timeout 300 python tools/bench_configs.py converge mega_cornell wf_cornell ao cornell512 4k synth > gpurun_out/configs_final.log 2>&1
cat gpurun_out/gpu_tests_final.log; cut -c1-300 gpurun_out/bench_final.log
