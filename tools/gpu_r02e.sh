#!/bin/bash
# round 2, visit e (2 GPUs): slab parity with the peer-memory epoch barrier / all-reduce, A/B against the NCCL barrier
o=gpurun_out; mkdir -p $o; tag=r02e
nvidia-smi -L | head -3
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -15 $o/${tag}_pytest.log | cut -c1-300
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 "$@"; }
echo "== default (peer epoch barrier + peer all-reduce)"; run > $o/${tag}_bench_2gpu.json 2> $o/${tag}_bench_2gpu.err; tail -c 1800 $o/${tag}_bench_2gpu.json; tail -3 $o/${tag}_bench_2gpu.err
echo "== PS3D_NO_PEER_SYNC=1"; PS3D_NO_PEER_SYNC=1 run > $o/${tag}_bench_2gpu_nccl.json 2> $o/${tag}_bench_2gpu_nccl.err; python -c "
import json; d=json.load(open('$o/${tag}_bench_2gpu_nccl.json')); print(d['ms_per_step'], d['parity'])"
echo "== 256^3 rk4"; run --grid 256 --stepper impl-diff-rk4 > $o/${tag}_bench_2gpu_rk4_256.json 2>/dev/null;  python -c "
import json; d=json.load(open('$o/${tag}_bench_2gpu_rk4_256.json')); print(d['ms_per_step'])"
