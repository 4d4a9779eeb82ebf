#!/bin/bash
# N-GPU measurement pass (one box): bench line + MSM sweep under torchrun.  usage: tools/measure_multi.sh <tag> <N>
tag=$1; N=$2; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 5 > $O/${tag}_bench_${N}gpu.json 2> $O/${tag}_bench_${N}gpu.err; echo "bench rc=$?"
timeout 900 $TR --master-port 29542 tools/msm_sweep.py --out $O/${tag}_msm_sweep_${N}gpu.json > /dev/null 2> $O/${tag}_sweep_${N}gpu.err; echo "sweep rc=$?"
if [ "$N" = 2 ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2; fi
tail -c 300 $O/${tag}_bench_${N}gpu.err
