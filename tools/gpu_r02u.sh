#!/bin/bash
# round 2, visit u (1 GPU): compute-sanitizer racecheck + memcheck of the rewritten column kernels (small grids, every stage)
o=gpurun_out; mkdir -p $o; tag=r02u
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tests/race_small.py 8x8x512 8x8x64 16x16x16 8x8x256 > $o/${tag}_racecheck.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|race_small ok|Error|hazard" $o/${tag}_racecheck.log | head -10; tail -4 $o/${tag}_racecheck.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tests/race_small.py 8x8x256 8x8x32 16x16x8 > $o/${tag}_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|race_small ok|Invalid" $o/${tag}_memcheck.log | head -10
