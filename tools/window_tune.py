"""MSM time per window width c for several N (calibrates pick_window's cost model)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
import bpperm_b200
be = bpperm_b200.Backend(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
be.set_stream(stream.cuda_stream)
d_out = torch.zeros(160, dtype=torch.uint8, device=dev)
for log_n in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["12", "14", "16", "18", "20"])]:
    n = 1 << log_n
    rs = np.random.RandomState(log_n)
    table = be.points_from_uniform(rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes())
    sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0F
    d_sc = torch.from_numpy(sc).to(dev)
    res = {}
    for c in [0] + list(range(8, 17)):
        be.set_window_bits(c)
        for _ in range(2):
            be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[c] = float(np.median(ts))
    be.set_window_bits(0)
    best = min((v, k) for k, v in res.items() if k)
    print(f"2^{log_n}: auto {res[0]:.3f} ms | " + " ".join(f"c={k}:{v:.3f}" for k, v in res.items() if k) + f" | best c={best[1]} {best[0]:.3f}", flush=True)
    table.free()
