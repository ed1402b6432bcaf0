#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv2d or many_tiles or operand_plane or final or graphed" 2>&1 | tail -8 > gpurun_out/r2_t5_pytest.log
tail -3 gpurun_out/r2_t5_pytest.log
for a in "2048 fp16 128 128 3 2 1 0" "2048 bf16x3 128 128 3 2 1 0" "1024 bf16x3 64 64 3 2 1 0" "1024 bf16x3 128 512 3 0 0 0" "1024 bf16x3 64 256 1 3 1 0" "2048 bf16x3 128 128 1 2 1 0"; do
  python tools/prof_conv.py $a
  RCN_TC_NMMA=1 python tools/prof_conv.py $a | sed 's/^/NMMA=1 /'
done > gpurun_out/r2_t5_prof.txt 2>&1
cat gpurun_out/r2_t5_prof.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t5_bench.json 2> gpurun_out/r2_t5_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t5_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['clocks'])"
tail -3 gpurun_out/r2_t5_bench.err
