#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t14_pytest.log
cat gpurun_out/r2_t14_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_t14_bench.json 2> gpurun_out/r2_t14_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t14_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value']);print(d['cpu_baseline']);print(d['symbols_mismatch_vs_oracle'])"
tail -2 gpurun_out/r2_t14_bench.err
python tools/trace_step.py 2048 bf16x3 forward > gpurun_out/r2_t14_trace.txt 2>&1
tail -1 gpurun_out/r2_t14_trace.txt
python tools/gma_sweep.py gpurun_out/r2_t14_gma_sweep.json > gpurun_out/r2_t14_gma.txt 2>&1; cat gpurun_out/r2_t14_gma.txt | tail -9
python tools/trace_step.py 256 bf16x3 gma > gpurun_out/r2_t14_gma_trace.txt 2>&1; head -24 gpurun_out/r2_t14_gma_trace.txt
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2_t14_ref.json 2> gpurun_out/r2_t14_ref.err; tail -c 900 gpurun_out/r2_t14_ref.json; tail -4 gpurun_out/r2_t14_ref.err
