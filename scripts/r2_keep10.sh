#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "kept or oracle or c_abi" 2>&1 | tail -2
timeout 300 scripts/r2_keep2.sh
RNNT_LIB_PATH=/root/repo/myrtlespeech_b200/lib/librnnt_prof.so python scripts/prof_frontend.py target 42 1 2>&1 | tail -5
