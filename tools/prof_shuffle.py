"""Two prove+verify steps of the 52-card batch (for ncu launch lists).  argv: batch, table window bits, mode"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench, bpperm_b200
be = bpperm_b200.Backend(0)
G = bpperm_b200.acproof
k = 52
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = sys.argv[3] if len(sys.argv) > 3 else "reference-fixed"
n, Q, m, WL, WR, WO, WV, c = bpperm_b200.weights.shuffle_circuit(k)
ng = G.next_pow2(n) if mode == "fixed" else n
rs = np.random.RandomState(4242)
pts = be.points_from_uniform(rs.randint(0, 256, size=(2 * ng + 2, 64), dtype=np.uint8).tobytes())
enc = be.compress_points(pts)
cir = G.Circuit(be, n, Q, m, WL, WR, WO, WV, c)
gens = G.Generators(be, enc[:32], enc[32:64], [enc[64 + 32 * i: 96 + 32 * i] for i in range(ng)],
                    [enc[64 + 32 * (ng + i): 96 + 32 * (ng + i)] for i in range(ng)], cb)
aL, aR, aO, gamma, v, seeds = bench.synth_shuffle_batch(k, B, 0)
batch = G.Batch(be, cir, gens, B, mode, b"test")
batch.upload_witness(aL, aR, aO, gamma, seeds)
batch.commit(v)
for _ in range(2):
    batch.prove()
    batch.verify(b"\x01" * 32)
be.synchronize()
print("accepted", batch.download_accept() == b"\x01" * B)
