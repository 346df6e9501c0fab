#!/bin/bash
mkdir -p gpurun_out
{
for v in "SOLB_WL_REGION_SLOTS=72" "SOLB_WL_REGION_SLOTS=48" "SOLB_WL_REGION_SLOTS=56" "SOLB_WL_REGION_SLOTS=64" "SOLB_WL_REGION_SLOTS=96" "SOLB_WL_REGION_SLOTS=128" "SOLB_WL_REGION_SLOTS=72 SOLB_WL_GEN_MIN=24" "SOLB_WL_REGION_SLOTS=96 SOLB_WL_FETCH_IDLE=16"; do
  for w in 8 4; do echo -n "$v : "; env $v timeout 120 python tools/tile_time.py $w 2>&1 | head -1; done
done
} > gpurun_out/r2_tile3.log 2>&1
cat gpurun_out/r2_tile3.log
