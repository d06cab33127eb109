#!/bin/bash
# round 2, visit i (1 GPU): column kernels with producer-side pre-processing (mirrored two-plane buffers) against the
# previous library on the same box; full -m gpu suite; ncu of the two column kernels
o=gpurun_out; mkdir -p $o; tag=r02i
( timeout 900 python -m pytest tests -m gpu -x -q ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-300
for v in tools/libps3d_cuda_prev.so ps3d_b200/libps3d_cuda.so; do
  echo "== $v"; PS3D_PROBE_LIB=$v timeout 200 python tools/gpu_probe.py 512 2>&1 | tee -a $o/${tag}_ab.log | cut -c1-200
done
PS3D_TRACE=1 timeout 200 python tools/gpu_probe.py 512 2>&1 | grep PS3D_TRACE > $o/${tag}_trace.log; head -12 $o/${tag}_trace.log | cut -c1-160
bash tools/gpu_ncu.sh $tag k_vor2vel_spec k_source_spec > /dev/null 2>&1
python tools/ncu_raw_summary.py $o/${tag}_raw_k_vor2vel_spec.csv | head -14
python tools/ncu_raw_summary.py $o/${tag}_raw_k_source_spec.csv | head -14
