#!/bin/bash
# Round-2 evidence: launch census of bench.py and smoke() under ncu, then one `ncu --set full` capture of the two
# persistent kernels.  Each ncu run follows a plain run of the same command that exited 0.
set -u
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
echo launch-rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_smoke.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_ncu_smoke.log 2>&1
echo smoke-rc=$?
python scripts/prof_one.py target > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bwd_mega|fwd_persist" -s 2 -c 2 -o gpurun_out/r2_prof_target \
    python scripts/prof_one.py target > gpurun_out/r2_ncu_full.log 2>&1
echo full-rc=$?
tail -3 gpurun_out/r2_ncu_full.log
grep -c . gpurun_out/r2_launches.csv gpurun_out/r2_launches_smoke.csv
