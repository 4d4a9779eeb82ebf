#!/bin/bash
# The 1-GPU measurement pass behind profiles/r2_*: bench line (driver's flags), reference arms, MSM sweep, batch sizes,
# the ncu launch list of the bench command and the --set full summaries.  usage: tools/measure_all.sh <tag> [quick]
tag=$1
O=gpurun_out
T0=$SECONDS; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${tag}_bench_full.json 2> $O/${tag}_bench_full.err; echo "bench rc=$? wall $((SECONDS - T0)) s"; tail -2 $O/${tag}_bench_full.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${tag}_bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$?"
timeout 600 python bench.py --impl reference --workload msm > $O/${tag}_cpu_msm_sweep.json 2>/dev/null; echo "cpu sweep rc=$?"
timeout 600 python tools/msm_sweep.py --out $O/${tag}_msm_sweep_1gpu.json > /dev/null 2>&1; echo "sweep rc=$?"
timeout 600 python tools/batch_sizes.py > $O/${tag}_shuffle_batch_sizes.jsonl 2>/dev/null; echo "batch sizes rc=$?"
timeout 600 python tools/ipa_fold_compare.py > /dev/null 2>&1; mv $O/ipa_fold_compare.json $O/${tag}_ipa_fold_compare.json
if [ "$2" != quick ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
  python tools/launch_summary.py $O/${tag}_launches_bench.csv 45 > $O/${tag}_launches_bench_summary.txt; gzip -f $O/${tag}_launches_bench.csv
  timeout 900 tools/ncu_full.sh $tag 2>&1 | tail -5
fi
du -sh $O
