#!/bin/bash
# fused LayerNorm + qkv embedding kernel; A/B of the LayerNorm fusions at step level; frame4k batch sizes through bench.py
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_ln_linear or fused_swin_mlp or tcm_blocks" 2>&1 | tail -25 > gpurun_out/r2_t27_pytest_fused.log
cat gpurun_out/r2_t27_pytest_fused.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t27_pytest.log
cat gpurun_out/r2_t27_pytest.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do set -- $cfg
  RCN_FUSED_LN_LINEAR=$1 RCN_FUSED_MLP_LN=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-frame 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('ln_linear=$1 mlp_ln=$2 step ms:', d['ms_per_step'], d['clocks']['sm_mhz'])"
done > gpurun_out/r2_t27_ab.txt 2>&1
cat gpurun_out/r2_t27_ab.txt
for b in 8 20 40; do RCN_FRAME_MAX_BATCH=$b timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);f=d['frame4k'];print('max_batch=$b', f.get('batch'), f.get('ms_per_frame'), f.get('value'), f.get('container_sha256'), f.get('error'))"
done > gpurun_out/r2_t27_frame_batch.txt 2>&1
cat gpurun_out/r2_t27_frame_batch.txt
