#!/bin/bash
# nested fork/join (attention-gate branches of SWAtten, conv / transformer halves of small-map ConvTransBlocks): correctness + A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t29_pytest.log
cat gpurun_out/r2_t29_pytest.log
for c in 1 0 1 0; do
  RCN_CONCURRENT_BRANCHES=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('concurrent=$c step ms:', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frame', d['frame4k'].get('ms_per_frame'), d['frame4k'].get('container_sha256'), 'decode', d['decode']['ms_per_tile'], d['clocks']['sm_mhz'])"
done > gpurun_out/r2_t29_ab.txt 2>&1
cat gpurun_out/r2_t29_ab.txt
