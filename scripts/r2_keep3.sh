#!/bin/bash
export RNNT_LIB_PATH=/root/repo/myrtlespeech_b200/lib/librnnt_prof.so
echo "== target keep wait counters"; python scripts/prof_mega.py target 42 2>&1 | tail -19
echo "== target dh epi phases"; python scripts/prof_dh_epi.py target 42 2>&1 | tail -10
