#!/bin/bash
# LayerNorm folded into the fused Swin-MLP kernel, 256-bit plane stores; frame4k batch-size sweep
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_swin_mlp or fused_ingest or tcm_blocks" 2>&1 | tail -25 > gpurun_out/r2_t26_pytest_fused.log
cat gpurun_out/r2_t26_pytest_fused.log
timeout 120 python tools/prof_ingest.py 2048 fused > gpurun_out/r2_t26_prof.txt 2>&1
timeout 120 python tools/prof_mlp.py 1024 fused >> gpurun_out/r2_t26_prof.txt 2>&1
cat gpurun_out/r2_t26_prof.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t26_pytest.log
cat gpurun_out/r2_t26_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t26_bench.json 2> gpurun_out/r2_t26_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t26_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value']); print(d['roofline_ingest']['ms_per_launch'])"
tail -2 gpurun_out/r2_t26_bench.err
RCN_FUSED_MLP_LN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-frame 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('LayerNorm as its own launch:', d['ms_per_step'])"
for b in 8 10 20 40; do RCN_FRAME_MAX_BATCH=$b timeout 300 python tools/frame_bench.py --reps 3 2>&1 | tail -1 | cut -c1-400 | sed "s/^/max_batch=$b /"; done > gpurun_out/r2_t26_frame_batch.txt 2>&1
cat gpurun_out/r2_t26_frame_batch.txt
