#!/bin/bash
# final state of the round: smoke(), full GPU suite, ncu launch list of one timed step, ncu --set full of the fused ingest kernel,
# bench line with the CPU reference leg
mkdir -p gpurun_out /tmp/ncu
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_t33_smoke.log 2>&1; tail -2 gpurun_out/r2_t33_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t33_pytest.log; cat gpurun_out/r2_t33_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file gpurun_out/r2_t33_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-frame > gpurun_out/r2_t33_ncu_bench.log 2>&1
wc -l gpurun_out/r2_t33_launches.csv
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${name}_source.csv.gz
  tail -1 gpurun_out/${name}_ncu.log; }
cap r2_t33_ingest ingest_kernel 2 python tools/prof_ingest.py 2048 fused
timeout 120 python tools/prof_ingest.py 2048 both > gpurun_out/r2_t33_prof.txt 2>&1
timeout 120 python tools/prof_mlp.py 1024 both >> gpurun_out/r2_t33_prof.txt 2>&1
cat gpurun_out/r2_t33_prof.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_t33_bench.json 2> gpurun_out/r2_t33_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_t33_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline_hbm']['ms_per_launch'],d['clocks'], d['gpu_launches']);print(d['frame4k']['ms_per_frame'], d['frame4k']['value'], d['decode']); print(d['cpu_baseline']['value'], d['symbols_mismatch_vs_oracle']); print(d['roofline_ingest'])"
tail -2 gpurun_out/r2_t33_bench.err
