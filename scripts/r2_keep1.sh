#!/bin/bash
timeout 200 python scripts/keep_vs_recompute.py small odd c2 2>&1 | grep -v Warning | tail -8
timeout 200 python scripts/keep_vs_recompute.py target 2>&1 | grep -v Warning | tail -4
timeout 300 scripts/r2_keep2.sh
timeout 200 scripts/r2_keep4.sh
