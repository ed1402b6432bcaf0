#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcm_blocks or final_forward or tcm_matches or small_ops" 2>&1 | tail -6 > gpurun_out/r2_t10_pytest_attn.log
cat gpurun_out/r2_t10_pytest_attn.log
for a in "1024 64 8 8 1" "1024 64 8 8 0" "512 64 16 8 1" "128 128 16 8 1" "256 64 32 8 1"; do
  python tools/prof_wmsa.py $a
  RCN_WMSA_FFMA=1 python tools/prof_wmsa.py $a | sed 's/^/FFMA /'
done > gpurun_out/r2_t10_wmsa.txt 2>&1
cat gpurun_out/r2_t10_wmsa.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_t10_pytest.log
cat gpurun_out/r2_t10_pytest.log
RCN_FRAME_TIMING=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t10_bench.json 2> gpurun_out/r2_t10_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t10_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'])"
grep "frame timing" gpurun_out/r2_t10_bench.err | tail -3
