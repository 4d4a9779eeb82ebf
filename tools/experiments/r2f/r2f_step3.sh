#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2f_pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2f_pytest3.log
BPPERM_LIB=bulletproof-perm_b200/variants/libbpperm_pp.so timeout 300 python tools/prof_round.py 52 reference-fixed 4096 16 2>/dev/null | head -1
for B in 2368 4096 4736 7104; do
  timeout 300 python tools/prof_round.py 52 reference-fixed $B 16 2>/dev/null | head -1
done
timeout 300 python tools/large_deck.py 4096 8 1 2>/dev/null | tail -1
