#!/bin/bash
# Final 1-GPU round of a session: parity tests, A/B of a switch, bench line, ncu launch list, ncu details (text only).
tag=${1:-rXX}; o=gpurun_out; mkdir -p $o
timeout 600 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $o/${tag}_pytest.log
for e in A=1 PS3D_NO_FUSED_UPDATE=1; do echo "== $e"; env $e timeout 200 python tools/gpu_probe.py 512 2>&1 | tee -a $o/${tag}_ab.log; done
timeout 300 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench exit $?"; cat $o/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py $o/${tag}_launches.csv $o/${tag}_launch_summary.csv "$tag: python bench.py --steps 1 --warmup 1 --no-cpu-baseline (Beltrami 512^3 cn2)" | tail -22
bash tools/gpu_ncu.sh $tag k_source_spec k_vor2vel_spec > /dev/null 2>&1
ls $o | grep $tag
