#!/bin/bash
# CUDA-graph replay of the ISP networks: test, throughput table (eager vs graph)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "isp_graph_replay or liteisp or isp_variants" 2>&1 | tail -6
timeout 600 python tools/liteisp_bench.py 256 1024 > gpurun_out/r2_t44_liteisp.jsonl 2> gpurun_out/r2_t44_liteisp.err; tail -2 gpurun_out/r2_t44_liteisp.err
python -c "
import json
for l in open('gpurun_out/r2_t44_liteisp.jsonl'):
    d=json.loads(l); print(d['model'], d['tile'], 'graph' if d['cuda_graph'] else 'eager', round(d['ms_per_tile'],2), 'ms', round(d['sensor_mp_per_s'],1), 'MP/s')"
