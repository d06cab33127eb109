#!/bin/bash
# round 2, visit w (1 GPU, final library): full -m gpu suite, smoke, bench line, ncu launch list of the bench command
o=gpurun_out; mkdir -p $o; tag=r02w
( timeout 900 python -m pytest tests -m gpu -x -q ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['step_roofline']['frac'], d['step_share_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $o/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py $o/${tag}_launches.csv $o/${tag}_launch_summary.csv "$tag: python bench.py --steps 1 --warmup 1 --no-cpu-baseline (Beltrami 512^3 cn2)" | head -12
