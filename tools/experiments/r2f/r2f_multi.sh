#!/bin/bash
# N-GPU pass: the bench line as the driver launches it, then end-to-end A/B legs (host threads sleeping in synchronise; four lanes)
N=$1; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 5 > $O/r2f_bench_${N}gpu.json 2> $O/r2f_bench_${N}gpu.err; echo "bench rc=$?"
if [ "$2" = ab ]; then
  BPP_BLOCKING_SYNC=1 timeout 600 $TR --master-port 29543 bench.py --gpus $N --steps 10 --warmup 5 --workload shuffle --no-cpu --no-extra --no-fixed > $O/r2f_shuffle_${N}gpu_blocking.json 2>/dev/null; echo "blocking rc=$?"
  BPP_LANES=4 timeout 600 $TR --master-port 29544 bench.py --gpus $N --steps 10 --warmup 5 --workload shuffle --no-cpu --no-extra --no-fixed > $O/r2f_shuffle_${N}gpu_lanes4.json 2>/dev/null; echo "lanes4 rc=$?"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2f_*_${N}gpu*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"]), "single", round(d["single_stream"]["value"]), "e2e", round(d["e2e"]["value"]), "serial", round(d["e2e"]["serial"]["value"]))
    except Exception as e: print(f, e)
PY
tail -c 300 $O/r2f_bench_${N}gpu.err
