#!/bin/bash
# round 2, visit v (1 GPU): named barriers per FFT group inside the column transform vs block-wide barriers; racecheck
o=gpurun_out; mkdir -p $o; tag=r02v
for v in tools/libps3d_cuda_prev.so ps3d_b200/libps3d_cuda.so tools/libps3d_cuda_prev.so ps3d_b200/libps3d_cuda.so; do
  echo "== $v"; PS3D_PROBE_LIB=$v timeout 200 python tools/gpu_probe.py 512 2>&1 | tee -a $o/${tag}_ab.log | cut -c1-160 | grep -v "fwd_\|inv_"
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tests/race_small.py 8x8x512 8x8x64 16x16x16 8x8x256 8x8x1024 > $o/${tag}_racecheck.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|race_small ok|hazard" $o/${tag}_racecheck.log | head -6
( timeout 600 python -m pytest tests -m gpu -x -q -k "white_noise or trajectory or buoyancy or config3 or known" ) 2>&1 | tail -2
