#!/bin/bash
# frame4k: tiles per graph replay (the sweep of session 27 did not reach the distributed path); frame_bench tool on the maintained pipeline
mkdir -p gpurun_out
for b in 8 10 20 40; do RCN_FRAME_MAX_BATCH=$b timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);f=d['frame4k'];print('max_batch=$b', f.get('batch'), f.get('ms_per_frame'), f.get('value'), f.get('container_sha256'), f.get('error'))"
done > gpurun_out/r2_t37_frame_batch.txt 2>&1
cat gpurun_out/r2_t37_frame_batch.txt
timeout 300 python tools/frame_bench.py --reps 3 2>&1 | tail -1 | cut -c1-500
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frame or tile_pipeline" 2>&1 | tail -3
