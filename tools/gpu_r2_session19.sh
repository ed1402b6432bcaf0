#!/bin/bash
# lean epilogue instantiations + compile-time halo issue loop: tests, the three reference launch shapes, bench line, call trace
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t19_pytest.log
cat gpurun_out/r2_t19_pytest.log
for a in "2048 bf16x3 128 128 1 2 1 0" "2048 fp16 128 128 3 2 1 0" "2048 bf16x3 128 128 3 2 1 0" "1024 bf16x3 64 256 1 3 1 0" "1024 bf16x3 64 64 3 2 1 0" "2048 bf16x3 4 128 3 0 1 0"; do
  python tools/prof_conv.py $a
done > gpurun_out/r2_t19_prof.txt 2>&1
cat gpurun_out/r2_t19_prof.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t19_bench.json 2> gpurun_out/r2_t19_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t19_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'])"
tail -2 gpurun_out/r2_t19_bench.err
python tools/trace_step.py 2048 bf16x3 forward > gpurun_out/r2_t19_trace.txt 2>&1
head -12 gpurun_out/r2_t19_trace.txt; tail -1 gpurun_out/r2_t19_trace.txt
