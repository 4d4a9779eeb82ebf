#!/bin/bash
# first pass over the digit-staged warp kernel and the fused inner-product round tail: parity, then A/B timings
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2f_pytest.log
for d in 1 0; do
  BPP_FB_DIGITS=$d timeout 300 python tools/prof_round.py 52 fixed 4096 16 2>/dev/null | head -1 > $O/r2f_fixed4096_digits$d.json; cat $O/r2f_fixed4096_digits$d.json
  BPP_FB_DIGITS=$d timeout 300 python tools/prof_round.py 52 reference-fixed 4096 16 2>/dev/null | head -1 > $O/r2f_reffixed4096_digits$d.json; cat $O/r2f_reffixed4096_digits$d.json
done
for cfg in "8 1" "8 0" "10 1" "11 1"; do
  set -- $cfg
  BPP_IPA_TAIL=$2 timeout 300 python tools/large_deck.py 4096 $1 1 2>/dev/null | tail -1 > $O/r2f_large_deck_c$1_tail$2.json; cat $O/r2f_large_deck_c$1_tail$2.json
done
