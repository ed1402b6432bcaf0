#!/bin/bash
# verification of the final library (host decoder with the 512-bucket table): smoke, full GPU suite, bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_t34_smoke.log 2>&1; tail -1 gpurun_out/r2_t34_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t34_pytest.log; cat gpurun_out/r2_t34_pytest.log
timeout 900 python bench.py > gpurun_out/r2_t34_bench.json 2> gpurun_out/r2_t34_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t34_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['clocks'], d['steps'], d['warmup']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'], d['decode']); print(d['cpu_baseline']['value'], d['symbols_mismatch_vs_oracle'])"
tail -2 gpurun_out/r2_t34_bench.err
