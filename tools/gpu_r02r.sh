#!/bin/bash
# round 2, visit r (1 GPU): two-deep streamed upload (copy of state k+1 also overlaps the decomposition of state k)
o=gpurun_out; mkdir -p $o; tag=r02r
( timeout 300 python -m pytest tests -m gpu -x -q -k "streamed" ) 2>&1 | tail -2
for k in 1 2; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_$k.json 2>/dev/null
python - $o/${tag}_$k.json <<'P'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
print(sys.argv[1], d['ms_per_step'], 'e2e streamed', e['streamed_ms_per_step'], 'serial', e['serial_ms_per_step'], e['value'])
P
done
