#!/bin/bash
# 2-GPU check of the bench (tile bench + frame4k scatter / gather over NCCL) with the multi-stream graphs, launched as the driver does
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_t30_bench_n2.json 2> gpurun_out/r2_t30_bench_n2.err
echo "rc=$?"; tail -3 gpurun_out/r2_t30_bench_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t30_bench_n2.json').read().strip().splitlines()[-1]);print(d['n_gpus'], d['ms_per_step'],d['value'],d['e2e']['value'],d['clocks']);print(d['frame4k'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_t30_ref_n2.json 2> gpurun_out/r2_t30_ref_n2.err
echo "ref rc=$?"; tail -c 600 gpurun_out/r2_t30_ref_n2.json
