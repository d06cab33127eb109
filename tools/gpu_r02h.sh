#!/bin/bash
# round 2, visit h (4 GPUs): 512^3 cn2 bench + in-situ kernel trace on 4 ranks
o=gpurun_out; mkdir -p $o; tag=r02h
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 "$@"; }
run > $o/${tag}_bench_4gpu.json 2> $o/${tag}_bench_4gpu.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02h_bench_4gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['step_share_ms'], d['nvlink'], d['e2e']['ms_per_step'], d['e2e']['serial_ms_per_step'], d['parity'])
P
PS3D_TRACE=1 run --steps 5 > /dev/null 2> $o/${tag}_trace_4gpu.err
grep PS3D_TRACE $o/${tag}_trace_4gpu.err | head -24 | cut -c1-160
PS3D_LINE_TMA=1 run --steps 5 > $o/${tag}_bench_4gpu_tma.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_4gpu_tma.json').read().strip().splitlines()[-1]); print('TMA=1', d['ms_per_step'])"
