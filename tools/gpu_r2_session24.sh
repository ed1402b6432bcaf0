#!/bin/bash
# ingest / Swin-MLP kernels with two epilogue warp groups per tile slot (640 threads): correctness, timing, bench
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_swin_mlp or fused_ingest or conditioning" 2>&1 | tail -25 > gpurun_out/r2_t24_pytest_fused.log
cat gpurun_out/r2_t24_pytest_fused.log
timeout 120 python tools/prof_ingest.py 2048 fused > gpurun_out/r2_t24_prof.txt 2>&1
cat gpurun_out/r2_t24_prof.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t24_bench.json 2> gpurun_out/r2_t24_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t24_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value']); print(d['roofline_ingest'])"
tail -2 gpurun_out/r2_t24_bench.err
timeout 300 python tools/trace_step.py 2048 bf16x3 forward 2>&1 | grep -n "mlp_fused\|ingest_fused\|step " > gpurun_out/r2_t24_trace_fused.txt
cat gpurun_out/r2_t24_trace_fused.txt
