#!/bin/bash
mkdir -p gpurun_out
{
for v in "SOLB_X=0" "SOLB_WL_FRAMES_IN_FLIGHT=1" "SOLB_WL_WARPS_PER_SM=16" "SOLB_WL_WARPS_PER_SM=20" "SOLB_WL_WARPS_PER_SM=24" "SOLB_WL_FRAMES_IN_FLIGHT=3"; do
  for w in 8 4; do echo -n "$v : "; env $v timeout 120 python tools/tile_time.py $w 2>&1 | head -1; done
done
} > gpurun_out/r2_tile2.log 2>&1
cat gpurun_out/r2_tile2.log
