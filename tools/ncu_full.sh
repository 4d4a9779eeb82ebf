#!/bin/bash
# ncu --set full captures (one GPU, never under a bench number): run AFTER the plain commands exited 0.
# usage: tools/ncu_full.sh <tag>   -> gpurun_out/<tag>_full_*.ncu-rep
set -x
tag=$1
NCU="ncu --set full --clock-control none --profile-from-start off"
$NCU -o gpurun_out/${tag}_full_fixed4096 -f python tools/prof_round.py 52 fixed 4096 16 > gpurun_out/${tag}_ncu_fixed.log 2>&1
BPP_GROUPS=1 ncu --set full --clock-control none -k regex:'k_bucket|k_digit|k_msm|k_node|k_window|k_sort|k_recode' -c 40 -o gpurun_out/${tag}_full_msm2p20 -f python tools/prof_msm.py > gpurun_out/${tag}_ncu_msm.log 2>&1
ls -la gpurun_out/*.ncu-rep
