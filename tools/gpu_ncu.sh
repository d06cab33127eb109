#!/bin/bash
# ncu --set full of the named kernels (one capture each, third matching launch); exports the details/raw/source
# pages as text on the box and drops the .ncu-rep (gpurun_out/ is capped at 64 MiB).
# usage: tools/gpu_ncu.sh <tag> <kernel-regex>...
tag=$1; shift
o=gpurun_out
mkdir -p $o
for k in "$@"; do
  n=$(echo $k | tr -c 'A-Za-z0-9_\n' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 2 -c 1 \
      -o /tmp/${tag}_$n -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_$n.log 2>&1
  ncu -i /tmp/${tag}_$n.ncu-rep --page details > $o/${tag}_details_$n.txt 2>&1
  ncu -i /tmp/${tag}_$n.ncu-rep --page raw --csv > $o/${tag}_raw_$n.csv 2>&1
  ncu -i /tmp/${tag}_$n.ncu-rep --page source --print-source cuda,sass --csv > $o/${tag}_source_$n.csv 2>&1
  gzip -f $o/${tag}_source_$n.csv
done
ls -la $o | tail -20
