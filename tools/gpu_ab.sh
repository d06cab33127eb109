#!/bin/bash
# A/B of library variants on one box: tools/gpu_ab.sh <tag> "<ENV=.. lib.so>" ...   (each item: optional env assignments + path)
tag=$1; shift
o=gpurun_out; mkdir -p $o
for item in "$@"; do
  echo "== $item" | tee -a $o/${tag}_ab.log
  lib=${item##* }; envs=${item% *}; [ "$envs" = "$item" ] && envs=""
  env $envs PS3D_PROBE_LIB=$lib timeout 200 python tools/gpu_probe.py 512 2>&1 | tee -a $o/${tag}_ab.log
done
