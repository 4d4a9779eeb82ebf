#!/bin/bash
# the driver's own sequence on one GPU: smoke(), then the bench line with the driver's flags and with no flags
O=gpurun_out
T0=$SECONDS; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1; echo "smoke wall $((SECONDS - T0)) s"
T0=$SECONDS; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2g_bench_full.json 2> $O/r2g_bench_full.err; echo "bench rc=$? wall $((SECONDS - T0)) s"
T0=$SECONDS; timeout 900 python bench.py > $O/r2g_bench_noflags.json 2>/dev/null; echo "bench (no flags) rc=$? wall $((SECONDS - T0)) s"
python tools/show_bench.py $O/r2g_bench_full.json; python tools/show_bench.py $O/r2g_bench_noflags.json
