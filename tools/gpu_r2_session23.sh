#!/bin/bash
# fused Swin MLP kernel: correctness (bounded), then the full suite and the bench line
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_swin_mlp or fused_ingest" 2>&1 | tail -25 > gpurun_out/r2_t23_pytest_mlp.log
cat gpurun_out/r2_t23_pytest_mlp.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t23_pytest.log
cat gpurun_out/r2_t23_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t23_bench.json 2> gpurun_out/r2_t23_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t23_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'])"
tail -2 gpurun_out/r2_t23_bench.err
RCN_FUSED_MLP=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-frame 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('unfused mlp:', d['ms_per_step'])"
timeout 300 python tools/trace_step.py 2048 bf16x3 forward > gpurun_out/r2_t23_trace.txt 2>&1
head -14 gpurun_out/r2_t23_trace.txt; tail -1 gpurun_out/r2_t23_trace.txt
