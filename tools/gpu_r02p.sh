#!/bin/bash
# round 2, visit p (1 GPU): streamed upload over one vs two copy streams (end-to-end loop of bench.py)
o=gpurun_out; mkdir -p $o; tag=r02p
show() { python - "$1" <<'P'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
print(sys.argv[1], d['ms_per_step'], 'e2e streamed', e['streamed_ms_per_step'], 'serial', e['serial_ms_per_step'])
P
}
for k in 1 2; do
PS3D_ONE_COPY_STREAM=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_one_$k.json 2>/dev/null; show $o/${tag}_one_$k.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_two_$k.json 2>/dev/null; show $o/${tag}_two_$k.json
done
