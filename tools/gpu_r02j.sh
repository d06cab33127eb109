#!/bin/bash
# round 2, visit j (1 GPU): column kernels with producer-side pre-processing, spills removed: A/B vs previous library, tests
o=gpurun_out; mkdir -p $o; tag=r02j
for v in tools/libps3d_cuda_prev.so ps3d_b200/libps3d_cuda.so; do
  echo "== $v"; PS3D_PROBE_LIB=$v timeout 200 python tools/gpu_probe.py 512 2>&1 | tee -a $o/${tag}_ab.log | cut -c1-200
done
( timeout 900 python -m pytest tests -m gpu -x -q ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-300
