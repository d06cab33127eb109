#!/bin/bash
# round 2, visit d: rotating TMA buffers; parity subset + same-box A/B
o=gpurun_out; mkdir -p $o; tag=r02d
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "not config3 and not config4" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -3 $o/${tag}_pytest.log | cut -c1-300
for v in 1 0 1 0; do
echo "== PS3D_TMA_ROT=$v" | tee -a $o/${tag}_ab.log
PS3D_TMA_ROT=$v timeout 300 python tools/gpu_probe.py 512 2>&1 | cut -c1-400 | tee -a $o/${tag}_ab.log
done
PS3D_TMA_ROT=1 timeout 300 python tools/gpu_probe.py 256 1024 2>&1 | cut -c1-300 | tee -a $o/${tag}_ab.log
