#!/bin/bash
# polyphase plane emission from stride-2 layers (CondNet2 / CondNet3 chains): tests + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conditioning or conv2d or operand_plane" 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t47_pytest.log; cat gpurun_out/r2_t47_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_t47_bench.json 2> gpurun_out/r2_t47_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t47_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['clocks'], d['gpu_launches']);print(d['frame4k']['ms_per_frame'], d['decode']['ms_per_tile'])"
tail -2 gpurun_out/r2_t47_bench.err
timeout 300 python tools/trace_step.py 2048 bf16x3 forward 2>&1 | grep -n "split_bf16_s2\|step \|by entry" | head -8 | cut -c1-300
