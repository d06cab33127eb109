#!/bin/bash
# round 2, visit l (4 GPUs): split-phase exchange protocol: slab parity at 2 and 4 ranks, 512^3 bench at 4 and 2 GPUs, A/B
o=gpurun_out; mkdir -p $o; tag=r02l
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-600
run() { np=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $np --steps 10 --warmup 3 "$@"; }
show() { python - "$1" <<'P'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['parity']['ok'], d['nvlink']['alltoalls_per_step'], d['nvlink']['scatter_sweep'])
P
}
run 4 > $o/${tag}_bench_4gpu.json 2> $o/${tag}_bench_4gpu.err; echo "bench4 exit $?"; show $o/${tag}_bench_4gpu.json
PS3D_NO_SPLIT_PHASE=1 run 4 > $o/${tag}_bench_4gpu_nosplit.json 2>/dev/null; show $o/${tag}_bench_4gpu_nosplit.json
run 2 > $o/${tag}_bench_2gpu.json 2> $o/${tag}_bench_2gpu.err; echo "bench2 exit $?"; show $o/${tag}_bench_2gpu.json
