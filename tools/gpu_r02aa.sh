#!/bin/bash
# round 2, visit aa (2 GPUs): 512^3 bench with the final defaults
o=gpurun_out; mkdir -p $o; tag=r02aa
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $o/${tag}_bench_2gpu.json 2> $o/${tag}_bench_2gpu.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02aa_bench_2gpu.json').read().strip().splitlines()[-1]); nv=d['nvlink']
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['upload'], d['parity']['ok'], {k: (round(v['GBs'],1) if isinstance(v, dict) and 'GBs' in v else None) for k, v in nv.items() if isinstance(v, dict)})
P
