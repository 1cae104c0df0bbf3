#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -k "wide" > gpurun_out/r2_wide2.log 2>&1; tail -3 gpurun_out/r2_wide2.log
python scripts/wide_v_bench.py 8,500,100,4096,1024 2>&1 | tail -2
python scripts/wide_v_bench.py 16,500,100,2048,1024 2>&1 | tail -2
python scripts/wide_v_bench.py 4,500,100,8192,1024 2>&1 | tail -2
