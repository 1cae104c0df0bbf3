#!/bin/bash
python scripts/active_frac.py target c2 c3 2>&1 | grep -v Warning | tail -9
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_hypothesis.py -x -q > gpurun_out/r2_pytest_gpu6.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu6.log
for w in target c2; do python bench.py --workload $w --steps 30 --warmup 10 --no-cpu-baseline --no-decode 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['clocks']['sm_mhz'])
"; done
