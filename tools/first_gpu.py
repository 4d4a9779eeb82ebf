"""First-contact script for the GPU box: IMAD peak, MSM phase timings (not a benchmark line)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bpperm_b200  # noqa: E402

be = bpperm_b200.Backend(0)
print("device", be.device_info())
for it in (2048, 8192):
    ops, ms = be.imad_peak(it)
    print(f"imad.wide peak: {ops/1e12:.3f} Tops/s in {ms:.3f} ms (iters={it})")
res = {}
for logn in (10, 14, 16, 18, 20):
    n = 1 << logn
    rs = np.random.RandomState(logn)
    t0 = time.time()
    table = be.points_from_uniform(rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes())
    t_up = time.time() - t0
    sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0F
    scb = sc.tobytes()
    be.set_profiling(True)
    out = None
    for c in ([0] if logn < 20 else [0, 14, 15, 16]):
        be.set_window_bits(c)
        for rep in range(3):
            t0 = time.time()
            out = be.vartime_multiscalar_mul(scb, table)
            dt = time.time() - t0
        ph = be.last_phase_ms()
        print(f"n=2^{logn} c={c} host-call {dt*1e3:.3f} ms phases(ms)={ {k: round(v,3) for k,v in ph.items()} } sum={sum(ph.values()):.3f} ops={be.last_op_counts()} upload {t_up:.2f}s out={out.hex()[:16]}")
        res[f"{logn}/{c}"] = ph
    be.set_window_bits(0)
    be.set_profiling(False)
    table.free()
json.dump(res, open("gpurun_out/first_gpu.json", "w"), indent=1)
