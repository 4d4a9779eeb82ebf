"""Stage timeline of one MSM (bpp_set_msm_trace) for a given size and number of window groups.
usage: python tools/msm_trace.py [log_n] [groups,groups,...]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
import bpperm_b200

be = bpperm_b200.Backend(0)
dev = torch.device("cuda", 0)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
groups = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "4"])]
n = 1 << log_n
rs = np.random.RandomState(log_n)
table = be.points_from_uniform(rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes())
sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
sc[:, 31] &= 0x0F
d_sc = torch.from_numpy(sc).to(dev)
d_out = torch.zeros(160, dtype=torch.uint8, device=dev)
for G in groups:
    be.set_msm_groups(G)
    for _ in range(3):
        be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
    be.synchronize()
    be.set_msm_trace(True)
    be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
    print(f"--- 2^{log_n} G={G}")
    print(be.msm_trace())
    be.set_msm_trace(False)
