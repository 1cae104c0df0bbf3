#!/bin/bash
# usage: scripts/exp_variants.sh v1 v2 ...  -> times bench.py (30 steps) with lib/librnnt_<v>.so
for v in "$@"; do
  RNNT_LIB_PATH=/root/repo/myrtlespeech_b200/lib/librnnt_$v.so timeout 300 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-decode 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'], d['clocks']['power_w_max'])
except Exception as e: print('$v', 'failed', e)
"
done
