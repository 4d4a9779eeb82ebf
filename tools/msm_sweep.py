#!/usr/bin/env python
"""BASELINE configs[4]: standalone ristretto255 MSM sweep 2^10..2^22 points on 1/2/4/8 GPUs (total N fixed,
points sharded contiguously over the ranks).  Every result is cross-checked on the GPU against a second window
decomposition (c = 13); the same seeded inputs are compared with the CPU restatement of dalek's
vartime_multiscalar_mul in tests/test_gpu_msm.py::test_sweep_inputs_match_oracle (2^10..2^16), which pins the
`result` bytes this tool prints.  `ms` is one bpp_msm_vartime_dev call (latency); on one GPU `submitted_ms` is the
per-MSM time of 16 MSMs submitted back to back with bpp_msm_submit_dev (two in flight: sustained throughput).

    python tools/msm_sweep.py --out gpurun_out/msm_sweep_1gpu.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/msm_sweep.py --out gpurun_out/msm_sweep_2gpu.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min", type=int, default=10)
    ap.add_argument("--max", type=int, default=22)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--out", default="gpurun_out/msm_sweep.json")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bpperm_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    be = bpperm_b200.Backend(local)
    stream = torch.cuda.current_stream(dev)
    be.set_stream(stream.cuda_stream)
    bpperm_b200.parallel.init_comm(be, world, rank, dev)   # the library's NCCL communicator (bpp_comm_init)
    rows = []
    for log_n in range(args.min, args.max + 1, args.step):
        n = 1 << log_n
        rs = np.random.RandomState(5000 + log_n)
        blobs = rs.randint(0, 256, size=(n, 64), dtype=np.uint8)
        sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
        sc[:, 31] &= 0x0F
        off, cnt = bpperm_b200.parallel.shard_bounds(n, world, rank)
        table = be.points_from_uniform(blobs[off:off + cnt].tobytes())
        d_sc = torch.from_numpy(sc[off:off + cnt].copy()).to(dev)
        sm = bpperm_b200.parallel.ShardedMsm(be, table, world, dev)

        def sync():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(3):
            sm.run(d_sc)
        sync()
        times = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync()
            e0.record(stream)
            sm.run(d_sc)
            e1.record(stream)
            sync()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            times.append(ms)
        res = bytes(sm.d_out[:32].cpu().numpy().tobytes())
        row = {"log_n": log_n, "n": n, "n_gpus": world, "ms": float(np.median(times)), "ms_min": float(min(times)),
               "points_per_s": n / (float(np.median(times)) * 1e-3), "result": res.hex()}
        # sustained form: 16 MSMs submitted back to back (two in flight per rank; sharded: the all-gather + sum of one
        # step on a side stream beside the next step's MSM), one wait; the result equals `res`
        k_sub = 16

        def burst():
            return sm.run_many([d_sc] * k_sub)

        burst()
        sync()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync()
            e0.record(stream)
            last = burst()
            e1.record(stream)
            sync()
            ms = e0.elapsed_time(e1) / k_sub
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            ts.append(ms)
        row["submitted_ms"] = float(np.median(ts))
        row["submitted_points_per_s"] = n / (row["submitted_ms"] * 1e-3)
        row["submitted_equal"] = bytes(last[:32].cpu().numpy().tobytes()) == res
        if world == 1:
            be.set_window_bits(13)
            alt = be.vartime_multiscalar_mul(sc.tobytes(), table)
            be.set_window_bits(0)
            row["check"] = "second window decomposition (c = 13)"
            row["equal"] = alt == res
        if rank == 0:
            print(json.dumps(row), flush=True)
        rows.append(row)
        sm.close()
        table.free()
        del d_sc, sm
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(rows, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
