"""MSM time for explicit window-group partitions (bpp_set_msm_partition), top group first.
usage: python tools/msm_partitions.py log_n "4,4,4,4;1,5,5,4,1;..." [trace]
Partitions that do not sum to the MSM's window count are skipped.  Results must equal the in-order bytes."""
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
log_n = int(sys.argv[1])
parts = [[int(x) for x in p.split(",")] for p in sys.argv[2].split(";")]
trace = len(sys.argv) > 3
n = 1 << log_n
rs = np.random.RandomState(log_n)
table = be.points_from_uniform(rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes())
sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
sc[:, 31] &= 0x0F
d_sc = torch.from_numpy(sc).to(dev)
d_out = torch.zeros(160, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
be.set_msm_groups(1)
be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
torch.cuda.synchronize()
want = bytes(d_out[:32].cpu().numpy())
for part in [[]] + parts:
    if part:
        be.set_msm_partition(part)
    for _ in range(3):
        be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
    torch.cuda.synchronize()
    assert bytes(d_out[:32].cpu().numpy()) == want, part
    ts = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"2^{log_n} partition {part or 'in-order'}: median {np.median(ts):.3f} ms, min {min(ts):.3f}", flush=True)
    if trace and part:
        be.set_msm_trace(True)
        be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
        print(be.msm_trace())
        be.set_msm_trace(False)
be.set_msm_partition([])
be.set_msm_groups(0)
