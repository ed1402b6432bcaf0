#!/bin/bash
# host decoder rewrite + e2e input prefetch: full suite, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t32_pytest.log
cat gpurun_out/r2_t32_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_t32_bench.json 2> gpurun_out/r2_t32_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t32_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks'], d['gpu_launches']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'], d['decode'])"
tail -2 gpurun_out/r2_t32_bench.err
