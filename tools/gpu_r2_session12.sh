#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t13_pytest.log
cat gpurun_out/r2_t13_pytest.log
for a in "1024 64 8 8 1" "512 64 16 8 1" "128 128 16 8 1" "256 64 32 8 1"; do python tools/prof_wmsa.py $a; done > gpurun_out/r2_t13_wmsa.txt 2>&1
cat gpurun_out/r2_t13_wmsa.txt
for a in "2048 fp16 128 128 3 2 1 0" "2048 bf16x3 128 128 3 2 1 0" "1024 bf16x3 64 64 3 2 1 0" "1024 bf16x3 64 256 1 3 1 0" "2048 bf16x3 128 128 1 2 1 0" "1024 bf16x3 256 64 1 0 1 1" "1024 bf16x3 64 192 1 0 0 0"; do
  python tools/prof_conv.py $a
done > gpurun_out/r2_t13_prof.txt 2>&1
cat gpurun_out/r2_t13_prof.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t13_bench.json 2> gpurun_out/r2_t13_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t13_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'])"
tail -2 gpurun_out/r2_t13_bench.err
python tools/trace_step.py 2048 bf16x3 forward > gpurun_out/r2_t13_trace.txt 2>&1
head -40 gpurun_out/r2_t13_trace.txt; tail -1 gpurun_out/r2_t13_trace.txt
