#!/bin/bash
# ingest kernel after the instruction-count / prefetch pass; A/B of a 16-epilogue-warp build of the conv engine
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_ingest or conditioning" 2>&1 | tail -5 > gpurun_out/r2_t22_pytest_ingest.log
cat gpurun_out/r2_t22_pytest_ingest.log
timeout 120 python tools/prof_ingest.py 2048 fused > gpurun_out/r2_t22_prof.txt 2>&1
cat gpurun_out/r2_t22_prof.txt
for v in "" tools/_variants/librcn_ep16.so; do
  for a in "2048 bf16x3 128 128 1 2 1 0" "2048 fp16 128 128 3 2 1 0" "2048 bf16x3 128 128 3 2 1 0" "1024 bf16x3 64 256 1 3 1 0" "1024 bf16x3 64 64 3 2 1 0" "1024 bf16x3 256 64 1 0 1 1" "1024 bf16x3 64 192 1 0 0 0"; do
    RCN_B200_LIB=$v timeout 100 python tools/prof_conv.py $a | sed "s|^|lib=${v:-default} |"
  done
done > gpurun_out/r2_t22_ep16.txt 2>&1
cat gpurun_out/r2_t22_ep16.txt
