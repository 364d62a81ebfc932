#!/bin/bash
# quick GPU check: parity tests + bench line summary (used during kernel iteration)
tag=${1:-q}
(timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -8) > gpurun_out/${tag}_tests.log 2>&1
(timeout 300 python bench.py --no-cpu-baseline ${BENCH_ARGS} 2>&1 | tail -3) > gpurun_out/${tag}_bench.log 2>&1
cat gpurun_out/${tag}_tests.log
python - <<PY
import json
l=[x for x in open("gpurun_out/${tag}_bench.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ("value","ms_per_step","kernels_ms_per_step","gpu_launches")}); print(d["e2e"])
else: print(open("gpurun_out/${tag}_bench.log").read())
PY
