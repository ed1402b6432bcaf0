#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
cap() {  # name, kernel regex, skip, cmd...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  gzip -f gpurun_out/${name}_source.csv
  tail -1 gpurun_out/${name}_ncu.log
}
cap r2_t11_mlp1 conv_tc 3 python tools/prof_conv.py 1024 bf16x3 64 256 1 3 1 0
cap r2_t11_tail conv_tc 3 python tools/prof_conv.py 2048 fp16 128 128 3 2 1 0
cap r2_t11_wmsa wmsa 2 python tools/prof_wmsa.py 1024 64 8 8 1
ls -la gpurun_out/r2_t11_*
