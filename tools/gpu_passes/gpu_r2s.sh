#!/bin/bash
# single GPU: full GPU suite, racecheck after the inactive-lane fix, default bench (pool publish leg, plans)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s_tests.log 2>&1
tail -n 5 gpurun_out/s_tests.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py quick > gpurun_out/s_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/s_racecheck.log
tail -n 3 gpurun_out/s_racecheck.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench_full.log 2>&1
SDRB_DEBUG_ONLY=filters timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-plans --no-zmq > gpurun_out/s_bench_filters.log 2>&1
tail -c 900 gpurun_out/s_bench_full.log
echo done
