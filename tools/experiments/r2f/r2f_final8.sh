#!/bin/bash
N=$1; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2g_bench_${N}gpu.json 2> $O/r2g_bench_${N}gpu.err; echo "bench rc=$?"
python tools/show_bench.py $O/r2g_bench_${N}gpu.json
