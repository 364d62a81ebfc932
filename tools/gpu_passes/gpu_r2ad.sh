#!/bin/bash
# single GPU: full GPU suite on the final tree + the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/ad_tests.log 2>&1
tail -n 5 gpurun_out/ad_tests.log
timeout 900 python bench.py > gpurun_out/ad_bench_default.log 2>&1
tail -c 400 gpurun_out/ad_bench_default.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ad_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/ad_smoke.log
tail -n 3 gpurun_out/ad_smoke.log
echo done
