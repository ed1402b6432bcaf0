#!/bin/bash
# fused ingest kernel: correctness first (bounded), then timing, full suite, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_ingest or conditioning" 2>&1 | tail -15 > gpurun_out/r2_t20_pytest_ingest.log
cat gpurun_out/r2_t20_pytest_ingest.log
timeout 120 python tools/prof_ingest.py 2048 both > gpurun_out/r2_t20_prof.txt 2>&1
cat gpurun_out/r2_t20_prof.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t20_pytest.log
cat gpurun_out/r2_t20_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t20_bench.json 2> gpurun_out/r2_t20_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t20_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'])"
tail -2 gpurun_out/r2_t20_bench.err
