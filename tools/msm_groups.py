"""MSM time per number of window groups G (bpp_set_msm_groups) for several N and window widths: calibrates
pick_groups / BPP_PIPELINE_MIN_POINTS.  Every G must give the bytes of G = 1 (asserted).
usage: python tools/msm_groups.py [log_n,log_n,...] [c,c,...] [out.json]"""
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
d_out = torch.zeros(160, dtype=torch.uint8, device=dev)
logs = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["16", "18", "19", "20", "21", "22"])]
cs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
out_path = sys.argv[3] if len(sys.argv) > 3 else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for log_n in logs:
    n = 1 << log_n
    rs = np.random.RandomState(log_n)
    table = be.points_from_uniform(rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes())
    sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0F
    d_sc = torch.from_numpy(sc).to(dev)
    for c in cs:
        be.set_window_bits(c)
        res, want = {}, None
        for G in [1, 2, 3, 4, 5, 6, 8, 0]:
            be.set_msm_groups(G)
            for _ in range(2):
                be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
            torch.cuda.synchronize()
            got = bytes(d_out[:32].cpu().numpy())
            if want is None:
                want = got
            assert got == want, (log_n, c, G, got.hex(), want.hex())
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                be.msm_dev(d_sc.data_ptr(), table, 0, n, d_out.data_ptr())
                e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[G] = float(np.median(ts))
        rows.append({"log_n": log_n, "c": c, "ms_by_groups": res})
        print(f"2^{log_n} c={c}: " + " ".join(f"G={k}:{v:.3f}" for k, v in res.items()), flush=True)
    be.set_window_bits(0)
    be.set_msm_groups(0)
    table.free()
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
