#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2_pytest_gpu5.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu5.log
for w in c2 target; do python bench.py --workload $w --steps 30 --warmup 10 --no-cpu-baseline --no-decode 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])
"; done
