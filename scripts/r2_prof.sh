#!/bin/bash
export RNNT_LIB_PATH=/root/repo/myrtlespeech_b200/lib/librnnt_prof.so
echo "== chunks, normal"; python scripts/prof_mega_chunks.py 0 2>&1 | tail -6
echo "== chunks, no TMA traffic after the first ring fill (dbg 256)"; python scripts/prof_mega_chunks.py 256 2>&1 | tail -6
echo "== wait counters"; python scripts/prof_mega.py target 2>&1 | tail -22
unset RNNT_LIB_PATH
echo "== kg sweep"; python scripts/kg_sweep.py target 3,2,4 2>&1 | tail -3
