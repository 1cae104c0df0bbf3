#!/bin/bash
python scripts/prof_one.py c2 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bwd_mega" -s 1 -c 1 -o gpurun_out/r2_prof_c2 python scripts/prof_one.py c2 > gpurun_out/r2_ncu_c2.log 2>&1
echo rc=$?; tail -2 gpurun_out/r2_ncu_c2.log
