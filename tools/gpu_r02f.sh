#!/bin/bash
# round 2, visit f (1 GPU): full -m gpu suite (incl. the buoyancy build), bench line, ncu --set full of the two column kernels
o=gpurun_out; mkdir -p $o; tag=r02f
( timeout 900 python -m pytest tests -m gpu -x -q ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -5 $o/${tag}_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench exit $?"
tail -c 2500 $o/${tag}_bench.json
bash tools/gpu_ncu.sh $tag k_vor2vel_spec k_source_spec > /dev/null 2>&1
ls -la $o | tail -8
