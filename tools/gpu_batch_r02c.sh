#!/bin/bash
# Round-2 final batch: speed check of the streaming kernel, full GPU tests, smoke, the three bench lines, racecheck of the final resident build.
cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 120 python tools/stream_syn_run.py 148 8 2 2>&1 | head -1 | tee $O/r02_stream_syn_final.log
timeout -s KILL 900 python -m pytest tests -q -m gpu --durations=3 2>&1 | tee $O/pytest_all_r2d.log | tail -8
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $O/r02_smoke.log | tail -6
timeout -s KILL 400 python bench.py 2>&1 | tail -1 > $O/r02_bench_1gpu_1024chains.json; cut -c1-300 $O/r02_bench_1gpu_1024chains.json
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $O/r02_bench_reference_arm.json; cut -c1-200 $O/r02_bench_reference_arm.json
timeout -s KILL 500 python bench.py --workload syn 2>&1 | tail -1 > $O/r02_bench_syn.json; cut -c1-300 $O/r02_bench_syn.json
timeout -s KILL 400 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitizer_run.py 4 20 0 toy > $O/r02_racecheck_resident_final_build_toy.log 2>&1; tail -3 $O/r02_racecheck_resident_final_build_toy.log
