#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t16_pytest.log
cat gpurun_out/r2_t16_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t16_bench.json 2> gpurun_out/r2_t16_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t16_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'])"
tail -2 gpurun_out/r2_t16_bench.err
python tools/trace_step.py 2048 bf16x3 forward > gpurun_out/r2_t16_trace.txt 2>&1
head -30 gpurun_out/r2_t16_trace.txt; tail -1 gpurun_out/r2_t16_trace.txt
