#!/bin/bash
# capture-only fork/join + ResidualUnit plane hand-over: full suite, bench, call trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t31_pytest.log
cat gpurun_out/r2_t31_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_t31_bench.json 2> gpurun_out/r2_t31_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t31_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks'], d['gpu_launches']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'], d['decode']['ms_per_tile']); print(d['roofline_ingest']['ms_per_launch'])"
tail -2 gpurun_out/r2_t31_bench.err
timeout 300 python tools/trace_step.py 2048 bf16x3 forward > gpurun_out/r2_t31_trace.txt 2>&1
head -8 gpurun_out/r2_t31_trace.txt; tail -1 gpurun_out/r2_t31_trace.txt
