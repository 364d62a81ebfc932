#!/bin/bash
# e2e (host-buffer) step time for different stream-group shapes of sdrb_bank_process_host
for g in "1,1,1,1,1,1,1,1" "1,1,1,1" "3,3,2,2,1,1" "4,4,3,2,1" "8,8,6,4,3,2,1" "6,5,4,3,2,1,1" "2,2,2,1,1"; do
  SDRB_HOST_GROUPS=$g timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$g', 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'dev_ms', round(d['ms_per_step'],3))"
done
