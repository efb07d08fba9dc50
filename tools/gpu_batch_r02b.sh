#!/bin/bash
# Round-2 batch B: A/B of the resident kernel (previous build vs advance_q fusion + Philox table), full GPU tests,
# racecheck on a warm-up transition that ends by the in-subtree U-turn break.
cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 200 python tools/ab_run.py _prev,default,_prev,default 148 100 50 2>&1 | tee $O/r02_ab_resident_fusion.log | tail -4
timeout -s KILL 900 python -m pytest tests -q -m gpu --durations=3 2>&1 | tee $O/pytest_all_r2c.log | tail -8
F=$(timeout -s KILL 120 python tools/sanitizer_run.py 16 8 0 find 2>&1 | tee $O/r02_find_partial_tree_warmup.log | grep FOUND | head -1)
echo "find (warm-up): $F"
OFF=$(echo "$F" | sed -n 's/.*offset=\([0-9]*\).*/\1/p'); IT=$(echo "$F" | sed -n 's/.*iters=\([0-9]*\).*/\1/p')
if [ -n "$OFF" ]; then
  timeout -s KILL 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitizer_run.py 1 $IT 0 offset=$OFF > $O/r02_racecheck_resident_subtree_break.log 2>&1; tail -3 $O/r02_racecheck_resident_subtree_break.log
  timeout -s KILL 400 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitizer_run.py 1 $IT 0 offset=$OFF > $O/r02_memcheck_resident_subtree_break.log 2>&1; tail -2 $O/r02_memcheck_resident_subtree_break.log
fi
