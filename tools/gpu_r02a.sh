#!/bin/bash
# round 2, visit a: full parity suite (new shapes, 256^3, 512^3 analytic, options), baseline bench, config-2 microbench,
# golden 512^3 series
o=gpurun_out; mkdir -p $o; tag=r02a
nproc > $o/${tag}_host.txt; free -g >> $o/${tag}_host.txt; nvidia-smi -L >> $o/${tag}_host.txt
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -30 $o/${tag}_pytest.log
timeout 300 python tools/microbench_64.py 64 > $o/${tag}_microbench_64.json 2> $o/${tag}_microbench_64.err; cat $o/${tag}_microbench_64.json
timeout 600 python tools/make_golden_512.py 512 40 $o/beltrami512_cn2_series.json 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench exit $?"
cat $o/${tag}_bench.json
timeout 400 python bench.py --steps 5 --warmup 3 --grid 256 --stepper impl-diff-rk4 --no-cpu-baseline > $o/${tag}_bench_rk4_256.json 2> $o/${tag}_bench_rk4_256.err; cat $o/${tag}_bench_rk4_256.json
