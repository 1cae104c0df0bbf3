#!/bin/bash
set -u
python scripts/prof_one.py target > gpurun_out/plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bwd_mega|fwd_persist" -s 2 -c 2 -o gpurun_out/r2_prof_keep \
    python scripts/prof_one.py target > gpurun_out/r2_ncu_keep.log 2>&1
echo full-rc=$?
tail -2 gpurun_out/r2_ncu_keep.log
