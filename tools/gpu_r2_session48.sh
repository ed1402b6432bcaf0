#!/bin/bash
# SWAtten plane hand-over (in_conv / SwinBlock -> first ResidualUnits): tests, bench, full call trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t48_pytest.log; cat gpurun_out/r2_t48_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_t48_bench.json 2> gpurun_out/r2_t48_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t48_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['clocks'], d['gpu_launches']);print(d['frame4k']['ms_per_frame'], d['decode']['ms_per_tile'])"
tail -2 gpurun_out/r2_t48_bench.err
RCN_TRACE_ROWS=400 timeout 300 python tools/trace_step.py 2048 bf16x3 forward > gpurun_out/r2_t48_trace_full.txt 2>&1
head -1 gpurun_out/r2_t48_trace_full.txt; grep "split_bf16" gpurun_out/r2_t48_trace_full.txt | head -30; tail -1 gpurun_out/r2_t48_trace_full.txt | cut -c1-300
