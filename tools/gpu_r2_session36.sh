#!/bin/bash
# fused kernels with all of a phase's TMEM loads requested at once: correctness, timing, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_ or conditioning or tcm_blocks" 2>&1 | tail -6 > gpurun_out/r2_t36_pytest_fused.log
cat gpurun_out/r2_t36_pytest_fused.log
timeout 120 python tools/prof_ingest.py 2048 fused > gpurun_out/r2_t36_prof.txt 2>&1
timeout 120 python tools/prof_mlp.py 1024 fused >> gpurun_out/r2_t36_prof.txt 2>&1
cat gpurun_out/r2_t36_prof.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_t36_bench.json 2> gpurun_out/r2_t36_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t36_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['decode']['ms_per_tile'], d['roofline_ingest']['ms_per_launch'])"
