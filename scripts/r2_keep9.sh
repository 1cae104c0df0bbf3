#!/bin/bash
timeout 200 python scripts/keep_vs_recompute.py small odd c2 target 2>&1 | grep -v Warning | tail -5 | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 300 scripts/r2_keep2.sh
