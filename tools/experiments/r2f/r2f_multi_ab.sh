#!/bin/bash
N=$1; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
A="bench.py --gpus $N --steps 12 --warmup 6 --workload shuffle --no-cpu --no-extra --no-fixed"
BPP_LANES=6 timeout 600 $TR --master-port 29545 $A > $O/r2f_shuffle_${N}gpu_lanes6.json 2>/dev/null; echo "lanes6 rc=$?"
BPP_LANES=6 BPP_BLOCKING_SYNC=1 timeout 600 $TR --master-port 29546 $A > $O/r2f_shuffle_${N}gpu_lanes6_blocking.json 2>/dev/null; echo "lanes6 blocking rc=$?"
BPP_LANES=8 timeout 600 $TR --master-port 29547 $A > $O/r2f_shuffle_${N}gpu_lanes8.json 2>/dev/null; echo "lanes8 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2f_shuffle_${N}gpu*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"]), "single", round(d["single_stream"]["value"]), "e2e", round(d["e2e"]["value"]), "serial", round(d["e2e"]["serial"]["value"]))
    except Exception as e: print(f, e)
PY
