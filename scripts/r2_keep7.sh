#!/bin/bash
export RNNT_LIB_PATH=/root/repo/myrtlespeech_b200/lib/librnnt_prof.so
echo "== keep"; python scripts/prof_persist.py target 2>&1 | tail -9
echo "== recompute"; RNNT_KEEP_ACTIVATIONS=0 python scripts/prof_persist.py target 2>&1 | tail -9
