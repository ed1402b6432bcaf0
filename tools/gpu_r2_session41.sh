#!/bin/bash
# 2-CTA clusters sharing the weight stages of the halo kernel (TMA multicast): correctness (bounded), A/B timing, full suite, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv2d or many_tiles or operand_plane" 2>&1 | tail -8 > gpurun_out/r2_t41_pytest_conv.log
cat gpurun_out/r2_t41_pytest_conv.log
for c in 1 0; do
  for a in "1024 bf16x3 64 64 3 2 1 0" "1024 bf16x3 64 128 3 2 1 0" "1024 bf16x3 128 128 3 2 1 0" "2048 bf16x3 128 128 3 2 1 0" "512 bf16x3 128 512 3 0 1 0"; do
    RCN_TC_CLUSTER=$c timeout 60 python tools/prof_conv.py $a | sed "s/^/cluster=$c /"
  done
done > gpurun_out/r2_t41_ab.txt 2>&1
cat gpurun_out/r2_t41_ab.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t41_pytest.log; cat gpurun_out/r2_t41_pytest.log
for c in 1 0; do RCN_TC_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-frame 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cluster=$c step ms:', d['ms_per_step'], d['clocks']['sm_mhz'])"; done > gpurun_out/r2_t41_bench_ab.txt 2>&1
cat gpurun_out/r2_t41_bench_ab.txt
