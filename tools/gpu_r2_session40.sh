#!/bin/bash
# triage of the 64->64 3x3 layer class (halo kernel, Ntile = 64): which role bounds it
mkdir -p gpurun_out
for dbg in 0 1 2 4 8 3 6 10; do
  for a in "1024 bf16x3 64 64 3 2 1 0" "1024 bf16x3 64 128 3 2 1 0"; do
    RCN_TC_DEBUG=$dbg timeout 60 python tools/prof_conv.py $a | sed "s/^/dbg=$dbg /"
  done
done > gpurun_out/r2_t40_triage.txt 2>&1
for st in 2 3 4; do RCN_TC_STAGES=$st timeout 60 python tools/prof_conv.py 1024 bf16x3 64 64 3 2 1 0 | sed "s/^/stages=$st /"; done >> gpurun_out/r2_t40_triage.txt 2>&1
cat gpurun_out/r2_t40_triage.txt
