#!/bin/bash
for w in small odd c2; do echo "== $w"; timeout 120 python scripts/keep_vs_recompute.py $w 2>&1 | grep -v Warning | tail -3 | cut -c1-300; done
