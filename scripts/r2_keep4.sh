#!/bin/bash
export RNNT_LIB_PATH=/root/repo/myrtlespeech_b200/lib/librnnt_prof.so
python scripts/prof_frontend.py target 42 1 2>&1 | tail -9
