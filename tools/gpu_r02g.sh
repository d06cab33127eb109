#!/bin/bash
# round 2, visit g (1 GPU): in-situ kernel trace of the 512^3 step (PS3D_TRACE), streamed end-to-end loop, new tests
o=gpurun_out; mkdir -p $o; tag=r02g
( timeout 600 python -m pytest tests -m gpu -x -q -k "streamed or buoyancy" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -3 $o/${tag}_pytest.log | cut -c1-300
PS3D_TRACE=1 timeout 300 python tools/gpu_probe.py 512 > $o/${tag}_trace.log 2>&1; grep -c PS3D_TRACE $o/${tag}_trace.log; head -40 $o/${tag}_trace.log | cut -c1-200
timeout 400 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
P
