#!/bin/bash
# round 2, visit n (1 GPU): full -m gpu suite, bench line, ncu launch list of the bench command, one-off 512^3 CPU step,
# reference arm, config 3 (256^3 rk4)
o=gpurun_out; mkdir -p $o; tag=r02n
( timeout 900 python -m pytest tests -m gpu -x -q ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02n_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['step_roofline']['frac'], d['step_share_ms'], d['e2e'], d['cpu_baseline']['value'])
P
timeout 300 python bench.py --grid 256 --stepper impl-diff-rk4 --steps 10 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_rk4_256.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02n_bench_rk4_256.json').read().strip().splitlines()[-1]); print('rk4 256', d['ms_per_step'], d['value'], d['step_roofline']['frac'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py $o/${tag}_launches.csv $o/${tag}_launch_summary.csv "$tag: python bench.py --steps 1 --warmup 1 --no-cpu-baseline (Beltrami 512^3 cn2)" | head -24
PS3D_TRACE=1 timeout 200 python tools/gpu_probe.py 512 2>&1 | grep PS3D_TRACE > $o/${tag}_trace.log; head -14 $o/${tag}_trace.log | cut -c1-150
timeout 600 python tools/cpu_512_once.py 512 1 $o/r02_cpu_512.json | cut -c1-400
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
