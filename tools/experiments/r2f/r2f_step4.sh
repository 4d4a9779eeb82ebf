#!/bin/bash
O=gpurun_out
for sm in 0 60000; do
  echo "== BPP_ACC_SMEM=$sm"; BPP_ACC_SMEM=$sm timeout 300 python tools/msm_stream.py 20 2>/dev/null | tail -2
done
for g in 1 2 3 4; do
  echo "== verify MSM groups=$g"; BPP_GROUPS=$g timeout 300 python tools/prof_round.py 52 reference-fixed 4096 16 2>/dev/null | head -1 | cut -c1-200
done
echo "== BPP_TR_WARP_MAX=5000"; BPP_TR_WARP_MAX=5000 timeout 300 python tools/prof_round.py 52 reference-fixed 4096 16 2>/dev/null | head -1 | cut -c1-200
