#!/bin/bash
# compute-sanitizer passes over the paths whose speed depends on internal streams, scratch slots and cross-stream events
# (VERDICT r1 next-round item 2d).  Logs -> gpurun_out/<tag>_sanitizer_*.log (summaries copied to profiles/).
# usage: tools/sanitize.sh <tag>
tag=$1
CS="compute-sanitizer --error-exitcode 7 --launch-timeout 0"
run() {   # name, tool, pytest args...
  local name=$1 tool=$2; shift 2
  echo "== $name ($tool): pytest $*" > gpurun_out/${tag}_sanitizer_${name}.log
  timeout 1500 $CS --tool $tool python -m pytest "$@" -x -q -m gpu -p no:cacheprovider >> gpurun_out/${tag}_sanitizer_${name}.log 2>&1
  echo "exit code $?" >> gpurun_out/${tag}_sanitizer_${name}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" gpurun_out/${tag}_sanitizer_${name}.log | tail -4
}
run msm_memcheck memcheck tests/test_gpu_msm.py -k "submitted or groups or hot or precomputed or batched"
run proofs_memcheck memcheck tests/test_gpu_acproof.py tests/test_gpu_ipa.py -k "small_decks or device_shuffle_witness or generator_fold or wire"
run proofs_racecheck racecheck tests/test_gpu_ipa.py tests/test_gpu_acproof.py -k "small_decks or device_shuffle_witness"
run msm_racecheck racecheck tests/test_gpu_msm.py -k "submitted or hot"
run ipa_synccheck synccheck tests/test_gpu_ipa.py -k "small_decks or larger_deck"
