#!/bin/bash
mkdir -p gpurun_out
for v in 0xFFFF 0xFFFE 0xFFFC 0xFFF0; do
  for a in "2048 bf16x3 128 128 1 2 1 0" "2048 fp16 128 128 3 2 1 0" "1024 bf16x3 64 256 1 3 1 0"; do
    RCN_B200_LIB=tools/_variants/librcn_$v.so python tools/prof_conv.py $a | sed "s/^/feat=$v /"
  done
done > gpurun_out/r2_t18_variants.txt 2>&1
cat gpurun_out/r2_t18_variants.txt
