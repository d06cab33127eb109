#!/bin/bash
# round 2, visit x (1 GPU): impl-diff-rk4 substeps riding on the source kernel: parity tests, 256^3 rk4 A/B, 512^3 cn2 check
o=gpurun_out; mkdir -p $o; tag=r02x
( timeout 600 python -m pytest tests -m gpu -x -q -k "rk4 or trajectory or resynchronised or separate_entry or config3 or buoyancy or options" ) 2>&1 | tail -3
for e in 0 1; do
  f=$o/${tag}_rk4_256_nofuse$e.json
  if [ $e = 1 ]; then PS3D_NO_FUSED_UPDATE=1 timeout 300 python bench.py --grid 256 --stepper impl-diff-rk4 --steps 10 --warmup 3 --no-cpu-baseline > $f 2>/dev/null
  else timeout 300 python bench.py --grid 256 --stepper impl-diff-rk4 --steps 10 --warmup 3 --no-cpu-baseline > $f 2>/dev/null; fi
  python -c "
import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'], d['step_roofline']['frac'])"
done
timeout 200 python tools/gpu_probe.py 512 2>&1 | grep -v "fwd_\|inv_" | cut -c1-160
