#!/bin/bash
# 4-GPU check of the bench line (tile bench + frame4k scatter / gather over NCCL), launched as the driver does
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_t35_bench_n4.json 2> gpurun_out/r2_t35_bench_n4.err
echo "rc=$?"; tail -3 gpurun_out/r2_t35_bench_n4.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t35_bench_n4.json').read().strip().splitlines()[-1]);print(d['n_gpus'], d['ms_per_step'],d['value'],d['e2e']['value'],d['clocks']);print(d['frame4k'])"
