#!/bin/bash
# ncu --set full captures of the kernels the roofline discussion names (one GPU, never under a bench number): run AFTER
# the plain commands exited 0.  A few launches per kernel name (-k regex, -s skip, -c count); every report is
# summarised on the box (tools/ncu_summary.py -> gpurun_out/<tag>_ncu_full_<name>.csv) and then deleted, except the
# top kernel's (kept for the source page), so that gpurun_out stays under its 64 MiB limit.
# usage: tools/ncu_full.sh <tag>
tag=$1
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # name, keep(0/1), regex, launch-skip, launch-count, command...
  local name=$1 keep=$2 re=$3 skip=$4 cnt=$5; shift 5
  local rep=gpurun_out/${tag}_full_${name}
  timeout 600 $NCU -k regex:"$re" -s $skip -c $cnt -o $rep -f "$@" > gpurun_out/${tag}_ncu_${name}.log 2>&1
  python tools/ncu_summary.py $rep.ncu-rep gpurun_out/${tag}_ncu_full_${name}.csv
  if [ "$keep" = 1 ]; then
    ncu -i $rep.ncu-rep --page source --csv > gpurun_out/${tag}_ncu_source_${name}.csv 2>/dev/null
    gzip -f gpurun_out/${tag}_ncu_source_${name}.csv
  fi
  rm -f $rep.ncu-rep
}
# 52-card `fixed` batch of 4096 (tools/prof_round.py: 3 warm-up steps, then 5 + 5 timed calls): skip the warm-up launches
cap fb_msm_warp 1 'k_fb_msm_warp' 30 1 python tools/prof_round.py 52 fixed 4096 16
cap ipa_round 0 'k_ipa_round|k_ipa_challenge' 20 4 python tools/prof_round.py 52 fixed 4096 16
cap acp_misc 0 'k_acp_decompress$|k_acp_dots|k_acp_vscal_fixed|k_compress_strided|k_pow_fill|k_acp_csr$|k_tr_verify|k_tr_vchunks|k_tr_weights' 30 12 python tools/prof_round.py 52 fixed 4096 16
cap large_deck_round 0 'k_ipa_lr_tail|k_ipa_round' 20 4 python tools/prof_round.py 4096 fixed 1 8
BPP_GROUPS=1 cap msm2p20 0 'k_bucket_accum|k_digit_scatter|k_digit_hist|k_bucket_fixup|k_msm_finish|k_node_merge' 20 14 python tools/prof_msm.py
du -sh gpurun_out
