#!/bin/bash
mkdir -p gpurun_out
K='conv2d or many_tiles or operand_plane'
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -6 > gpurun_out/r2_t7_pytest.log
cat gpurun_out/r2_t7_pytest.log
for a in "2048 fp16 128 128 3 2 1 0" "2048 bf16x3 128 128 3 2 1 0" "1024 bf16x3 64 64 3 2 1 0" "1024 bf16x3 128 512 3 0 0 0" "1024 bf16x3 64 256 1 3 1 0" "2048 bf16x3 128 128 1 2 1 0" "1024 bf16x3 256 64 1 0 1 1" "1024 bf16x3 64 192 1 0 0 0"; do
  python tools/prof_conv.py $a
  RCN_TC_RV=0 python tools/prof_conv.py $a | sed 's/^/RV=0 /'
done > gpurun_out/r2_t7_prof.txt 2>&1
cat gpurun_out/r2_t7_prof.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_t7_pytest_all.log
cat gpurun_out/r2_t7_pytest_all.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t7_bench.json 2> gpurun_out/r2_t7_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t7_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']);print(d['decode'])"
tail -3 gpurun_out/r2_t7_bench.err
