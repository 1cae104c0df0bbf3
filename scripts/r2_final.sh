#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_final.log
python bench.py > gpurun_out/r2_bench_final.log 2>&1; tail -1 gpurun_out/r2_bench_final.log | cut -c1-300
python bench.py --workload c2 --no-decode > gpurun_out/r2_bench_c2.log 2>&1; tail -1 gpurun_out/r2_bench_c2.log | cut -c1-300
python bench.py --workload c3 --no-decode --no-cpu-baseline > gpurun_out/r2_bench_c3.log 2>&1; tail -1 gpurun_out/r2_bench_c3.log | cut -c1-300
python bench.py --ragged --no-decode --no-cpu-baseline > gpurun_out/r2_bench_ragged.log 2>&1; tail -1 gpurun_out/r2_bench_ragged.log | cut -c1-300
python bench.py --impl reference > gpurun_out/r2_bench_ref.log 2>&1; tail -1 gpurun_out/r2_bench_ref.log | cut -c1-400
