#!/bin/bash
# round 2, visit z (1 GPU): last full -m gpu run and smoke on the committed library
o=gpurun_out; mkdir -p $o; tag=r02z
( timeout 900 python -m pytest tests -m gpu -x -q ) > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $o/${tag}_pytest.log
tail -4 $o/${tag}_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python tools/gpu_probe.py 512 2>&1 | grep -v "fwd_\|inv_" | cut -c1-170
