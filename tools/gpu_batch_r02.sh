#!/bin/bash
# One GPU-box batch of round 2: sanitizer on a transition that ends by the in-subtree U-turn break, ncu captures of the
# sampling-phase launches of both kernels, launch lists of bench.py, and the full GPU test suite.  Outputs -> gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
F=$(timeout -s KILL 120 python tools/sanitizer_run.py 64 0 6 adapted find 2>&1 | tee $O/r02_find_partial_tree.log | grep FOUND | head -1)
echo "find: $F"
OFF=$(echo "$F" | sed -n 's/.*offset=\([0-9]*\).*/\1/p'); IT=$(echo "$F" | sed -n 's/.*iters=\([0-9]*\).*/\1/p')
if [ -n "$OFF" ]; then
  timeout -s KILL 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitizer_run.py 1 0 $IT adapted offset=$OFF > $O/r02_racecheck_resident_subtree_break.log 2>&1; tail -3 $O/r02_racecheck_resident_subtree_break.log
  timeout -s KILL 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitizer_run.py 1 0 $IT adapted offset=$OFF > $O/r02_memcheck_resident_subtree_break.log 2>&1; tail -2 $O/r02_memcheck_resident_subtree_break.log
  timeout -s KILL 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitizer_run.py 1 0 $IT adapted stream offset=$OFF > $O/r02_racecheck_stream_subtree_break.log 2>&1; tail -3 $O/r02_racecheck_stream_subtree_break.log
fi
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:potus_nuts_kernel -s 1 -c 1 -o $O/r02_resident_sampling python tools/profile_run.py 148 150 2 > $O/r02_ncu_resident_sampling.log 2>&1; tail -2 $O/r02_ncu_resident_sampling.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:potus_stream_kernel -s 1 -c 1 -o $O/r02_stream_sampling python tools/stream_syn_run.py 148 8 1 > $O/r02_ncu_stream_sampling.log 2>&1; tail -2 $O/r02_ncu_stream_sampling.log
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_2016.csv python bench.py --steps 1 --warmup 1 --chains 148 --iter-warmup 60 --iter-sampling 50 --no-cpu-baseline > $O/r02_launches_2016.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_syn.csv python bench.py --workload syn --steps 1 --warmup 1 --iter-warmup 10 --iter-sampling 4 --no-cpu-baseline > $O/r02_launches_syn.log 2>&1
timeout -s KILL 900 python -m pytest tests -q -m gpu --durations=5 2>&1 | tee $O/pytest_all_r2b.log | tail -12
