#!/bin/bash
# final tree: memcheck on the paths added last (condition UNet planes, fused kernels), smoke, full GPU suite, bench line with the CPU leg
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conditioning or small_ops or fused_ or many_tiles" > gpurun_out/r2_t46_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_t46_memcheck.log; tail -4 gpurun_out/r2_t46_memcheck.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_t46_smoke.log 2>&1; tail -1 gpurun_out/r2_t46_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t46_pytest.log; cat gpurun_out/r2_t46_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_t46_bench.json 2> gpurun_out/r2_t46_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t46_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['frac'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'], d['decode']['ms_per_tile']); print(d['cpu_baseline']['value'], d['symbols_mismatch_vs_oracle']['mismatching'], d['roofline_ingest']['ms_per_launch'])"
tail -2 gpurun_out/r2_t46_bench.err
