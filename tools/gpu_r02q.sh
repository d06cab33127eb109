#!/bin/bash
# round 2, visit q (4 GPUs): four vs two rotating receive buffers; slab parity incl. the buoyancy build; peer-copy ceilings
o=gpurun_out; mkdir -p $o; tag=r02q
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-600
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline; }
show() { python - "$1" <<'P'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); nv=d['nvlink']
print(sys.argv[1], d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['parity']['ok'], {k: (round(v['GBs'],1) if isinstance(v, dict) and 'GBs' in v else None) for k, v in nv.items() if isinstance(v, dict)})
P
}
PS3D_ROT_BUFS=2 run > $o/${tag}_4gpu_rot2.json 2>/dev/null; show $o/${tag}_4gpu_rot2.json
run > $o/${tag}_4gpu_rot4.json 2> $o/${tag}_4gpu_rot4.err; show $o/${tag}_4gpu_rot4.json
PS3D_ROT_BUFS=2 run > $o/${tag}_4gpu_rot2b.json 2>/dev/null; show $o/${tag}_4gpu_rot2b.json
run > $o/${tag}_4gpu_rot4b.json 2>/dev/null; show $o/${tag}_4gpu_rot4b.json
