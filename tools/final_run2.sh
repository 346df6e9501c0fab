#!/bin/bash
# session-3 secondary evidence: new knob tests, configs[2] in full, configs[4] on one GPU, cornell / AO lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "async_polls or voted_traversal" 2>&1 | tail -5
timeout 600 python tools/bench_configs.py converge mega_cornell wf_cornell ao cornell512 4k > gpurun_out/configs_s3.log 2>&1
timeout 900 python tools/bench_configs.py synth >> gpurun_out/configs_s3.log 2>&1
grep config gpurun_out/configs_s3.log | cut -c1-260
