#!/bin/bash
# round 2, visit m (4 GPUs): blocks per SM of the persistent scatter sweeps x split-phase protocol (512^3 cn2)
o=gpurun_out; mkdir -p $o; tag=r02m
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline; }
show() { python - "$1" <<'P'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['parity']['ok'], d['nvlink']['scatter_sweep']['ms'])
P
}
for ctas in 1 2; do for ns in 0 1; do
  f=$o/${tag}_4gpu_ctas${ctas}_nosplit${ns}.json
  if [ $ns = 1 ]; then PS3D_P2P_CTAS=$ctas PS3D_NO_SPLIT_PHASE=1 run > $f 2>/dev/null; else PS3D_P2P_CTAS=$ctas run > $f 2>/dev/null; fi
  show $f
done; done
