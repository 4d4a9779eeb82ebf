import sys
import numpy as np
sys.path.insert(0, ".")
import bpperm_b200
be = bpperm_b200.Backend(0)
import os
be.set_msm_groups(int(os.environ.get("BPP_GROUPS", "0")))   # 1 = in-order kernels (whole-MSM launches) for ncu
n = 1 << 20
rs = np.random.RandomState(20)
table = be.points_from_uniform(rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes())
sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
sc[:, 31] &= 0x0F
for _ in range(3):
    out = be.vartime_multiscalar_mul(sc.tobytes(), table)
print(out.hex())
