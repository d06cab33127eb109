#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of the column + line kernels.
# usage: tools/gpu_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-rXX}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -3 $o/${tag}_pytest.log
if [ -n "$RACE" ]; then
  timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python tests/race_small.py 8x8x512 8x8x64 16x16x16 > $o/${tag}_racecheck.log 2>&1; echo "racecheck exit $?"
  tail -4 $o/${tag}_racecheck.log
fi
for v in $VARIANTS; do
  echo "== variant $v"; PS3D_PROBE_LIB=$v timeout 200 python tools/gpu_probe.py 512 2>&1 | tee -a $o/${tag}_variants.log
done
timeout 300 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench exit $?"
cat $o/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py $o/${tag}_launches.csv $o/${tag}_launch_summary.csv "$tag: python bench.py --steps 1 --warmup 1 --no-cpu-baseline (Beltrami 512^3 cn2)"
for k in k_vor2vel_spec k_source_spec k_line_fwd k_line_inv; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 \
      -o $o/${tag}_full_$k -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_full_$k.log 2>&1
  ncu -i $o/${tag}_full_$k.ncu-rep --page details > $o/${tag}_details_$k.txt 2>&1
done
ls -la $o | tail -12
