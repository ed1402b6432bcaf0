#!/bin/bash
mkdir -p gpurun_out
K='conv2d or many_tiles or operand_plane'
echo "== halo, base_offset = (start>>7)&7" > gpurun_out/r2_t6_pytest.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -6 >> gpurun_out/r2_t6_pytest.log
echo "== halo, base_offset = 0 (RCN_TC_DEBUG=64)" >> gpurun_out/r2_t6_pytest.log
RCN_TC_DEBUG=64 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -6 >> gpurun_out/r2_t6_pytest.log
cat gpurun_out/r2_t6_pytest.log
for a in "2048 fp16 128 128 3 2 1 0" "2048 bf16x3 128 128 3 2 1 0" "1024 bf16x3 64 64 3 2 1 0" "1024 bf16x3 128 512 3 0 0 0" "128 bf16x3 576 224 3 3 0 0"; do
  python tools/prof_conv.py $a
  RCN_TC_DEBUG=64 python tools/prof_conv.py $a | sed 's/^/BO0 /'
  RCN_TC_HALO=0 python tools/prof_conv.py $a | sed 's/^/HALO=0 /'
done > gpurun_out/r2_t6_prof.txt 2>&1
cat gpurun_out/r2_t6_prof.txt
