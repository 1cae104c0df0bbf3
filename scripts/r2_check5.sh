#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -k "occupancy or opcheck or live_graphs" > gpurun_out/r2_pytest_gpu7.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu7.log
python bench.py --no-cpu-baseline --no-decode 2>&1 | tail -1 > gpurun_out/r2_bench_skip.log; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_skip.log').read()); print(d['ms_per_step'], d['value'], 'every tile:', d['ms_per_step_every_tile'], d['value_every_tile'], d['backward_tiles'], {k:v['ms_per_step'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['roofline']['hw_frac'], d['frac_of_bf16_peak'], d['clocks']['sm_mhz'])"
