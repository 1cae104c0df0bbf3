#!/bin/bash
for k in 0 1; do for w in target c2; do RNNT_KEEP_ACTIVATIONS=$k python bench.py --workload $w --steps 20 --warmup 8 --no-cpu-baseline --no-decode 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('keep=$k $w', d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])
"; done; done
