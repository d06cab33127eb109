#!/bin/bash
# Multi-GPU lines on one 8-GPU box: tools/gpu_scale.sh <tag>
tag=${1:-rXX}; o=gpurun_out; mkdir -p $o
run() { # n grid steps warmup
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1 + $2 % 97)) \
      bench.py --gpus $1 --grid $2 --steps $3 --warmup $4 2> $o/${tag}_bench_$2_$1gpu.err | grep "^{" > $o/${tag}_bench_$2_$1gpu.json
  python - <<PY
import json
try:
    d = json.load(open("$o/${tag}_bench_$2_$1gpu.json"))
    print("$2^3 x$1:", round(d["ms_per_step"], 2), "ms/step", "%.3e" % d["value"], "step_roofline", round(d["step_roofline"]["frac"], 3), "e2e %.3e" % d["e2e"]["value"])
except Exception as e:
    print("$2^3 x$1: failed", e)
PY
}
timeout 500 python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -3
run 8 512 10 3
run 4 512 10 3
run 8 1024 5 2
tail -3 $o/${tag}_bench_1024_8gpu.err
