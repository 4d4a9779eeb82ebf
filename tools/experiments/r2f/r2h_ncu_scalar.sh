#!/bin/bash
# --set full summaries of the scalar kernels rewritten last (warp-per-dot, Montgomery-form factors)
rep=gpurun_out/r2h_full_scalar
timeout 100 ncu --set full --clock-control none --import-source on -k regex:'k_acp_dots_warp|k_acp_vscal|k_acp_final$' -s 8 -c 4 -o $rep -f python tools/prof_round.py 52 reference-fixed 4096 16 > gpurun_out/r2h_ncu_scalar.log 2>&1
python tools/ncu_summary.py $rep.ncu-rep gpurun_out/r2h_ncu_full_scalar.csv
rm -f $rep.ncu-rep
