#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_acproof.py tests/test_gpu_ipa.py tests/test_gpu_transcript.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
for nc in "" 1; do
  echo "== BPP_NO_CRIT=$nc"
  BPP_NO_CRIT=$nc BPP_ACP_TRACE=1 timeout 300 python tools/prof_round.py 52 reference-fixed 4096 16 2>&1 | grep "acp trace\|prove_ms" | tail -3 | cut -c1-330
  BPP_NO_CRIT=$nc timeout 300 python tools/prof_round.py 52 fixed 4096 16 2>&1 | grep "prove_ms" | cut -c1-200
done
