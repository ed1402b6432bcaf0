#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t3_pytest.log
tail -3 gpurun_out/r2_t3_pytest.log
python tools/precision_table.py 512 2048 > gpurun_out/r2_t3_precision.md 2> gpurun_out/r2_t3_precision.err
cat gpurun_out/r2_t3_precision.md; tail -3 gpurun_out/r2_t3_precision.err
python tools/liteisp_bench.py 256 1024 > gpurun_out/r2_t3_liteisp.jsonl 2> gpurun_out/r2_t3_liteisp.err
cat gpurun_out/r2_t3_liteisp.jsonl; tail -3 gpurun_out/r2_t3_liteisp.err
# the dominant kernel as the step launches it: fp16 single pass, operand planes in / out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 3 -c 1 -f -o gpurun_out/r2_t3_conv_tail_fp16 python tools/prof_conv.py 2048 fp16 128 128 3 2 1 0 > gpurun_out/r2_t3_ncu_conv.log 2>&1
tail -2 gpurun_out/r2_t3_ncu_conv.log
# launch list of one timed step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file gpurun_out/r2_t3_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_t3_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r2_t3_ncu_bench.log
