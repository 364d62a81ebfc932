#!/bin/bash
# usage: tools/variant_sweep.sh TAG "ENV1=.. ENV2=.." "ENV.." ...   -- one short kernel-only bench per environment setting
tag=$1; shift
i=0
for envs in "$@"; do
  (env $envs timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 30 2>&1 | tail -1) > gpurun_out/${tag}_$i.log
  python - "$envs" gpurun_out/${tag}_$i.log <<'PY'
import json, sys
l = [x for x in open(sys.argv[2]) if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print(sys.argv[1], "|", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["kernels_ms_per_step"].items()})
else:
    print(sys.argv[1], "| FAILED", open(sys.argv[2]).read()[-400:])
PY
  i=$((i+1))
done
