#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv2d or many_tiles or operand_plane" 2>&1 | tail -5 > gpurun_out/r2_t15_pytest_conv.log
cat gpurun_out/r2_t15_pytest_conv.log
for a in "2048 bf16x3 4 128 3 0 1 0" "2048 bf16x3 32 16 3 1 0 0" "2048 bf16x3 16 64 3 1 1 0" "2048 bf16x3 4 16 3 1 0 0" "2048 fp16 128 16 3 0 0 0" "1024 bf16x3 64 32 3 1 0 0"; do
  python tools/prof_conv.py $a
  RCN_TC_HALO=64 python tools/prof_conv.py $a | sed 's/^/HALO=64 /'
done > gpurun_out/r2_t15_prof.txt 2>&1
cat gpurun_out/r2_t15_prof.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_t15_pytest.log
cat gpurun_out/r2_t15_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t15_bench.json 2> gpurun_out/r2_t15_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t15_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'])"
tail -2 gpurun_out/r2_t15_bench.err
