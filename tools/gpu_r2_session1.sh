#!/bin/bash
# round 2, GPU session 1: full GPU suite (incl. the T=2048 timed-configuration parity test), engine lever sizing, sanitizer evidence
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2_t1_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_t1_bench_bf16x3.json 2> gpurun_out/r2_t1_bench_bf16x3.err
python bench.py --steps 5 --warmup 3 --engine bf16 --no-cpu-baseline > gpurun_out/r2_t1_bench_bf16.json 2> gpurun_out/r2_t1_bench_bf16.err
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv2d or many_tiles or operand_plane" > gpurun_out/r2_t1_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_t1_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_tiles or epilogues_and_views" > gpurun_out/r2_t1_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_t1_racecheck.log
tail -5 gpurun_out/r2_t1_pytest.log
tail -3 gpurun_out/r2_t1_memcheck.log
tail -3 gpurun_out/r2_t1_racecheck.log
