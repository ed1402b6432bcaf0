#!/bin/bash
# ncu --set full of the fused ingest kernel (CONV variant) and of the lsc-only variant
mkdir -p gpurun_out /tmp/ncu
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${name}_source.csv.gz
  tail -1 gpurun_out/${name}_ncu.log; }
cap r2_t21_ingest ingest_kernel 2 python tools/prof_ingest.py 2048 fused
