"""Sustained MSM throughput: K independent MSMs back to back, (a) one call at a time (bpp_msm_vartime_dev: the
caller's stream joins every MSM) and (b) submitted (bpp_msm_submit_dev + one bpp_msm_wait: two in flight).
Every submitted result must equal the joined one.
usage: python tools/msm_stream.py [log_n] ["part;part;..."] [K] [out.json]"""
import json
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
import bpperm_b200

be = bpperm_b200.Backend(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
be.set_stream(stream.cuda_stream)
import os
be.set_msm_tile(int(os.environ.get("BPP_TILE", "0")))
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
parts = [[int(x) for x in p.split(",")] for p in sys.argv[2].split(";")] if len(sys.argv) > 2 and sys.argv[2] else [[]]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out_path = sys.argv[4] if len(sys.argv) > 4 else None
n = 1 << log_n
rs = np.random.RandomState(log_n)
table = be.points_from_uniform(rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes())
sets = []
for i in range(4):
    sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0F
    sets.append(torch.from_numpy(sc).to(dev))
outs = torch.zeros(K, 160, dtype=torch.uint8, device=dev)
rows = []


def run(submit):
    for i in range(K):
        (be.msm_submit_dev if submit else be.msm_dev)(sets[i % 4].data_ptr(), table, 0, n, outs[i].data_ptr())
    if submit:
        be.msm_wait()


def timed(submit):
    run(submit)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(submit)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / K)
    return float(np.median(ts))


be.set_msm_groups(1)
t_inorder = timed(False)
want = outs[:, :32].cpu().numpy().copy()
t_inorder_sub = timed(True)
print(f"2^{log_n} in-order, joined: {t_inorder:.3f} ms/MSM (submitted in order: {t_inorder_sub:.3f})", flush=True)
rows.append({"log_n": log_n, "partition": "in-order", "joined_ms": t_inorder})
be.set_msm_groups(0)
for part in parts:
    be.set_msm_partition(part)
    outs.zero_()
    tj = timed(False)
    assert (outs[:, :32].cpu().numpy() == want).all(), ("joined", part)
    outs.zero_()
    ts = timed(True)
    assert (outs[:, :32].cpu().numpy() == want).all(), ("submitted", part)
    print(f"2^{log_n} partition {part or 'auto'}: joined {tj:.3f} ms/MSM, submitted {ts:.3f} ms/MSM "
          f"({n / ts / 1e3:.1f} M points/s)", flush=True)
    rows.append({"log_n": log_n, "partition": part or "auto", "joined_ms": tj, "submitted_ms": ts})
be.set_msm_partition([])
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
