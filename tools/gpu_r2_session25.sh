#!/bin/bash
# round-2 final evidence: sanitizer on the fused kernels, ncu launch list of one timed step, ncu --set full of the two fused kernels,
# full GPU suite, bench line with the CPU reference leg
mkdir -p gpurun_out /tmp/ncu
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_ingest or fused_swin_mlp or conv2d or many_tiles or operand_plane" > gpurun_out/r2_t25_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_t25_memcheck.log; tail -4 gpurun_out/r2_t25_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_ingest or fused_swin_mlp or many_tiles" > gpurun_out/r2_t25_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_t25_racecheck.log; tail -4 gpurun_out/r2_t25_racecheck.log
timeout 120 python tools/prof_mlp.py 1024 both > gpurun_out/r2_t25_prof_mlp.txt 2>&1; cat gpurun_out/r2_t25_prof_mlp.txt
timeout 120 python tools/prof_ingest.py 2048 both > gpurun_out/r2_t25_prof_ingest.txt 2>&1; cat gpurun_out/r2_t25_prof_ingest.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file gpurun_out/r2_t25_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-frame > gpurun_out/r2_t25_ncu_bench.log 2>&1
tail -c 200 gpurun_out/r2_t25_ncu_bench.log; wc -l gpurun_out/r2_t25_launches.csv
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${name}_source.csv.gz
  tail -1 gpurun_out/${name}_ncu.log; }
cap r2_t25_ingest ingest_kernel 2 python tools/prof_ingest.py 2048 fused
cap r2_t25_mlp mlp_fused 2 python tools/prof_mlp.py 1024 fused
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t25_pytest.log; cat gpurun_out/r2_t25_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_t25_bench.json 2> gpurun_out/r2_t25_bench.err
tail -c 1500 gpurun_out/r2_t25_bench.json
