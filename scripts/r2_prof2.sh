#!/bin/bash
export RNNT_LIB_PATH=/root/repo/myrtlespeech_b200/lib/librnnt_prof.so
echo "== c2 chunks"; WL=c2 P=63 python scripts/prof_mega_chunks.py 0 2>&1 | tail -6
echo "== c2 wait counters"; python scripts/prof_mega.py c2 63 2>&1 | tail -21
echo "== target chunks"; python scripts/prof_mega_chunks.py 0 2>&1 | tail -6
