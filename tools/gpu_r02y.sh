#!/bin/bash
# round 2, visit y (1 GPU, final library): full -m gpu suite, smoke, bench line, ncu launch list of the bench command
o=gpurun_out; mkdir -p $o; tag=r02y
( timeout 900 python -m pytest tests -m gpu -x -q ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['step_roofline']['frac'], d['step_share_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
P
