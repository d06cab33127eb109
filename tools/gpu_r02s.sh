#!/bin/bash
# round 2, visit s (final) (8 GPUs): slab parity at 2/4/8 ranks, 512^3 bench at 8 GPUs with the 1024^3 config5 block, in-situ trace
o=gpurun_out; mkdir -p $o; tag=r02s
nvidia-smi -L | wc -l
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-400
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 "$@"; }
run > $o/${tag}_bench_8gpu.json 2> $o/${tag}_bench_8gpu.err; echo "bench exit $?"; tail -3 $o/${tag}_bench_8gpu.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02s_bench_8gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['step_share_ms'], d['nvlink'], d['e2e']['ms_per_step'], d['e2e']['serial_ms_per_step'], d['parity'])
print(json.dumps(d.get('config5'))[:1500])
P
