"""Shuffle-path tuning probe: per-step time and commitment-MSM kernel time for several table window widths.
argv: batch, comma-separated window widths, mode (reference-fixed | fixed)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench, bpperm_b200
be = bpperm_b200.Backend(0)
G = bpperm_b200.acproof
k, B = 52, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n, Q, m, WL, WR, WO, WV, c = bpperm_b200.weights.shuffle_circuit(k)
mode = sys.argv[3] if len(sys.argv) > 3 else "reference-fixed"
ng = G.next_pow2(n) if mode == "fixed" else n
rs = np.random.RandomState(4242)
pts = be.points_from_uniform(rs.randint(0, 256, size=(2 * ng + 2, 64), dtype=np.uint8).tobytes())
enc = be.compress_points(pts)
cir = G.Circuit(be, n, Q, m, WL, WR, WO, WV, c)
aL, aR, aO, gamma, v, seeds = bench.synth_shuffle_batch(k, B, 0)
ref = None
for cb in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["8", "10", "12"])]:
    t0 = time.time()
    gens = G.Generators(be, enc[:32], enc[32:64], [enc[64 + 32 * i: 96 + 32 * i] for i in range(ng)],
                        [enc[64 + 32 * (ng + i): 96 + 32 * (ng + i)] for i in range(ng)], cb)
    tb = time.time() - t0
    batch = G.Batch(be, cir, gens, B, mode)
    batch.upload_witness(aL, aR, aO, gamma, seeds)
    batch.commit(v)
    for _ in range(2):
        batch.prove(); batch.verify(b"\x01" * 32)
    be.synchronize()
    t0 = time.time()
    R = 3
    for _ in range(R):
        batch.prove(); be.synchronize()
    tp = (time.time() - t0) / R
    t0 = time.time()
    for _ in range(R):
        batch.verify(b"\x01" * 32); be.synchronize()
    tv = (time.time() - t0) / R
    ok = batch.download_accept() == b"\x01" * B
    pr = batch.download_proofs()
    if ref is None:
        ref = pr
    fb = batch.time_commit_msm(3)
    print(f"c={cb:2d} table build {tb:.2f}s prove {tp*1e3:.2f} ms verify {tv*1e3:.2f} ms -> {B/(tp+tv):.0f} proofs/s; fb kernel {fb[0]:.3f} ms ok={ok} same_bytes={pr==ref}")
    batch.free(); gens.free()
