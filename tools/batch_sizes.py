"""BASELINE configs[1]/[3]: 52-card prove+verify throughput at batch sizes 1, 64, 4096, both
protocol modes, and verification of a 4096-proof batch in which 1 % of the proofs are corrupted (fall-back path)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench, bpperm_b200

be = bpperm_b200.Backend(0)
G = bpperm_b200.acproof
k = 52
rows = []
for mode in ("reference-fixed", "fixed"):
    n, Q, m, WL, WR, WO, WV, c = bpperm_b200.weights.shuffle_circuit(k)
    ng = G.next_pow2(n) if mode == "fixed" else n
    rs = np.random.RandomState(4242)
    pts = be.points_from_uniform(rs.randint(0, 256, size=(2 * ng + 2, 64), dtype=np.uint8).tobytes())
    enc = be.compress_points(pts)
    cir = G.Circuit(be, n, Q, m, WL, WR, WO, WV, c)
    gens = G.Generators(be, enc[:32], enc[32:64], [enc[64 + 32 * i: 96 + 32 * i] for i in range(ng)],
                        [enc[64 + 32 * (ng + i): 96 + 32 * (ng + i)] for i in range(ng)], 16)
    for B in (1, 64, 4096):
        aL, aR, aO, gamma, v, seeds = bench.synth_shuffle_batch(k, B, 0)
        batch = G.Batch(be, cir, gens, B, mode, b"test")
        batch.upload_witness(aL, aR, aO, gamma, seeds)
        Vc = batch.commit(v)
        for _ in range(2):
            batch.prove(); batch.verify(b"\x01" * 32)
        be.synchronize()
        R = 5
        t0 = time.time()
        for _ in range(R):
            batch.prove(); be.synchronize()
        tp = (time.time() - t0) / R
        t0 = time.time()
        for _ in range(R):
            batch.verify(b"\x01" * 32); be.synchronize()
        tv = (time.time() - t0) / R
        ok = batch.download_accept() == b"\x01" * B
        row = {"mode": mode, "batch": B, "prove_ms": tp * 1e3, "verify_ms": tv * 1e3, "proofs_per_s": B / (tp + tv),
               "proof_bytes": batch.proof_len, "all_accepted": ok}
        if B == 4096:   # 1 % corrupted: the combined check fails, the per-proof kernels decide
            proofs = bytearray(batch.download_proofs())
            bad = list(range(7, B, 100))
            for p in bad:
                proofs[p * batch.proof_len + 40] ^= 1
            batch.upload_proofs(bytes(proofs), Vc)
            batch.verify(b"\x01" * 32); be.synchronize()
            t0 = time.time()
            for _ in range(R):
                batch.verify(b"\x01" * 32); be.synchronize()
            tvb = (time.time() - t0) / R
            acc = batch.download_accept()
            row["verify_ms_1pct_corrupted"] = tvb * 1e3
            row["corrupted_rejected_only"] = [i for i in range(B) if acc[i] == 0] == bad
        print(json.dumps(row), flush=True)
        rows.append(row)
        batch.free()
    gens.free(); cir.free()
json.dump(rows, open("gpurun_out/batch_sizes.json", "w"), indent=1)
