#!/bin/bash
# round-end evidence run on a GPU box: knob sweep, bench lines, ncu launch list + one full capture of the dominant kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash tools/sweep.sh "SOLB_TRI_WEIGHT=2" "SOLB_FETCH_IDLE=6" "SOLB_FETCH_IDLE=12" "SOLB_SHADE_CTAS_PER_SM_OVERLAP=2" > gpurun_out/sweep_s3.log 2>&1
python bench.py > gpurun_out/bench_s3.log 2> gpurun_out/bench_s3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_s3.log 2> gpurun_out/bench_reference_s3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01_s3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_s3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 30 -c 1 -f -o gpurun_out/prof_wf_trace_r01_s3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_s3.log 2>&1
tail -2 gpurun_out/sweep_s3.log; cat gpurun_out/bench_s3.log | cut -c1-400
