#!/bin/bash
# final evidence: sanitizer on the final library, ncu launch list of one timed step, ncu --set full of the dominant kernel, bench line
mkdir -p gpurun_out /tmp/ncu
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv2d or many_tiles or operand_plane or tcm_blocks or small_ops or conditioning" > gpurun_out/r2_t17_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_t17_memcheck.log; tail -4 gpurun_out/r2_t17_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_tiles or epilogues_and_views or tcm_blocks" > gpurun_out/r2_t17_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_t17_racecheck.log; tail -4 gpurun_out/r2_t17_racecheck.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file gpurun_out/r2_t17_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-frame > gpurun_out/r2_t17_ncu_bench.log 2>&1
tail -c 200 gpurun_out/r2_t17_ncu_bench.log; wc -l gpurun_out/r2_t17_launches.csv
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${name}_source.csv.gz
  tail -1 gpurun_out/${name}_ncu.log; }
cap r2_t17_tail conv_tc 3 python tools/prof_conv.py 2048 fp16 128 128 3 2 1 0
cap r2_t17_tail3 conv_tc 3 python tools/prof_conv.py 2048 bf16x3 128 128 3 2 1 0
cap r2_t17_k1 conv_tc 3 python tools/prof_conv.py 2048 bf16x3 128 128 1 2 1 0
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_t17_bench.json 2> gpurun_out/r2_t17_bench.err
tail -c 600 gpurun_out/r2_t17_bench.json
