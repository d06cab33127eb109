#!/bin/bash
# round 2, visit o (1 GPU): non-inlined transform in the column kernels (instruction-cache footprint) vs inlined
o=gpurun_out; mkdir -p $o; tag=r02o
for v in tools/libps3d_cuda_prev.so ps3d_b200/libps3d_cuda.so tools/libps3d_cuda_prev.so ps3d_b200/libps3d_cuda.so; do
  echo "== $v"; PS3D_PROBE_LIB=$v timeout 200 python tools/gpu_probe.py 512 2>&1 | tee -a $o/${tag}_ab.log | cut -c1-200
done
( timeout 600 python -m pytest tests -m gpu -x -q -k "white_noise or trajectory or buoyancy or config3" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -3 $o/${tag}_pytest.log | cut -c1-300
