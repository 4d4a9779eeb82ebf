#!/bin/bash
# launch lists (warm caches: --cache-control none) of the three proof configurations, block-size variants of the warp
# kernel, and a --set full capture of k_fb_msm_warp_d
O=gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv"
timeout 300 $NCU --log-file $O/r2f_launches_large_deck.csv python tools/prof_round.py 4096 fixed 1 8 > $O/r2f_ld.log 2>&1; echo "ld rc=$?"; head -1 $O/r2f_ld.log
timeout 300 $NCU --log-file $O/r2f_launches_fixed4096.csv python tools/prof_round.py 52 fixed 4096 16 > $O/r2f_f.log 2>&1; echo "fixed rc=$?"; head -1 $O/r2f_f.log
timeout 300 $NCU --log-file $O/r2f_launches_reffixed4096.csv python tools/prof_round.py 52 reference-fixed 4096 16 > $O/r2f_rf.log 2>&1; echo "reffixed rc=$?"; head -1 $O/r2f_rf.log
for v in t64 t256; do
  BPPERM_LIB=bulletproof-perm_b200/variants/libbpperm_$v.so timeout 300 python tools/prof_round.py 52 reference-fixed 4096 16 2>/dev/null | head -1 > $O/r2f_reffixed4096_$v.json; cat $O/r2f_reffixed4096_$v.json
done
rep=$O/r2f_full_fb_msm_warp_d
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fb_msm_warp_d -s 30 -c 1 -o $rep -f python tools/prof_round.py 52 fixed 4096 16 > $O/r2f_ncu_fb_msm_warp_d.log 2>&1
python tools/ncu_summary.py $rep.ncu-rep $O/r2f_ncu_full_fb_msm_warp_d.csv
ncu -i $rep.ncu-rep --page source --csv > $O/r2f_ncu_source_fb_msm_warp_d.csv 2>/dev/null; gzip -f $O/r2f_ncu_source_fb_msm_warp_d.csv
rm -f $rep.ncu-rep
ls -la $O | grep r2f
