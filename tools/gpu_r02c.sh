#!/bin/bash
# round 2, visit c: adapt restructured (fused strain + char. vorticity, kept x-transformed velocity), TMA sweeps with
# 16-z and 8-z tiles; parity subset, same-box A/B, ncu of the TMA sweeps
o=gpurun_out; mkdir -p $o; tag=r02c
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_options.py tests/test_golden.py -m gpu -x -q -k "not config3 and not config4" ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -5 $o/${tag}_pytest.log | cut -c1-300
echo "== default (TMA 16-z)" | tee -a $o/${tag}_ab.log
timeout 300 python tools/gpu_probe.py 512 2>&1 | cut -c1-400 | tee -a $o/${tag}_ab.log
echo "== PS3D_TMA_ZC=8" | tee -a $o/${tag}_ab.log
PS3D_TMA_ZC=8 timeout 300 python tools/gpu_probe.py 512 2>&1 | cut -c1-400 | tee -a $o/${tag}_ab.log
echo "== PS3D_NO_KEEP_VELX=1" | tee -a $o/${tag}_ab.log
PS3D_NO_KEEP_VELX=1 timeout 300 python tools/gpu_probe.py 512 2>&1 | head -1 | cut -c1-400 | tee -a $o/${tag}_ab.log
echo "== PS3D_LINE_TMA=0" | tee -a $o/${tag}_ab.log
PS3D_LINE_TMA=0 timeout 300 python tools/gpu_probe.py 512 2>&1 | head -1 | cut -c1-400 | tee -a $o/${tag}_ab.log
PS3D_TMA_ZC=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "operators_white_noise or trajectory" 2>&1 | tail -2
for k in k_line_tma; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 2 \
      -o $o/${tag}_full_$k -f python tools/gpu_probe.py 512 > $o/${tag}_ncu_full_$k.log 2>&1
  ncu -i $o/${tag}_full_$k.ncu-rep --page details > $o/${tag}_details_$k.txt 2>&1
  ncu -i $o/${tag}_full_$k.ncu-rep --page raw --csv > $o/${tag}_raw_$k.csv 2>&1
done
ls -la $o | tail -5
