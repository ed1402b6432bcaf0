#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t2_pytest.log
tail -3 gpurun_out/r2_t2_pytest.log
python tools/precision_table.py 512 2048 > gpurun_out/r2_t2_precision.md 2> gpurun_out/r2_t2_precision.err
cat gpurun_out/r2_t2_precision.md; tail -3 gpurun_out/r2_t2_precision.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_t2_bench.json 2> gpurun_out/r2_t2_bench.err
tail -c 1500 gpurun_out/r2_t2_bench.json; tail -3 gpurun_out/r2_t2_bench.err
python tools/trace_step.py 2048 bf16x3 forward > gpurun_out/r2_t2_trace.txt 2>&1
head -5 gpurun_out/r2_t2_trace.txt
