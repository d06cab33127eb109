#!/bin/bash
# round 2, visit b: TMA-staged line sweeps -- parity subset, then same-box A/B against the register-staged sweeps
o=gpurun_out; mkdir -p $o; tag=r02b
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_stafft_lit.py tests/test_golden.py -m gpu -x -q -k "not config3" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -5 $o/${tag}_pytest.log | cut -c1-300
for v in 0 1; do
  echo "== PS3D_LINE_TMA=$v" | tee -a $o/${tag}_ab.log
  PS3D_LINE_TMA=$v timeout 300 python tools/gpu_probe.py 512 256 2>&1 | cut -c1-400 | tee -a $o/${tag}_ab.log
done
PS3D_LINE_TMA=1 timeout 200 python tools/gpu_probe.py 1024x 2>&1 | tail -3
