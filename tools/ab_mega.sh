#!/bin/bash
# A/B of the warp-voted traversal in the one-pixel-per-lane kernels (SOLB_MEGA_VOTE) on a GPU box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for v in 0 1; do
SOLB_MEGA_VOTE=$v timeout 120 python tools/bench_configs.py ao mega_tunnel --frames 6 2>&1 | grep config
done
