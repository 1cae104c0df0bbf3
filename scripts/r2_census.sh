#!/bin/bash
# Launch census of bench.py and smoke() under ncu (each after a plain run of the same command that exited 0).
set -u
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
echo launch-rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_smoke.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_ncu_smoke.log 2>&1
echo smoke-rc=$?
tail -2 gpurun_out/plain2.log
grep -c . gpurun_out/r2_launches.csv gpurun_out/r2_launches_smoke.csv
