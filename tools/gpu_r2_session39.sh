#!/bin/bash
# ncu --set full of the 64->64 3x3 layer class (halo kernel, Ntile = 64) at 1024^2
mkdir -p gpurun_out /tmp/ncu
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${name}_source.csv.gz
  tail -1 gpurun_out/${name}_ncu.log; }
cap r2_t39_c64 conv_tc 3 python tools/prof_conv.py 1024 bf16x3 64 64 3 2 1 0
cap r2_t39_c64to128 conv_tc 3 python tools/prof_conv.py 1024 bf16x3 64 128 3 2 1 0
