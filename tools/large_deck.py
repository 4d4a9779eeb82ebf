"""BASELINE configs[2]: large-deck shuffle (k = 4096 cards -> n = 8192, 2^14 generators, 13 IPA rounds), `fixed`
mode, single proof latency and small-batch throughput on one GPU.  argv: k, window bits, batch sizes (comma)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bpperm_b200

be = bpperm_b200.Backend(0)
G = bpperm_b200.acproof
W = bpperm_b200.weights
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
batches = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1", "8"])]
n, Q, m, WL, WR, WO, WV, c = W.shuffle_circuit(k)
ng = G.next_pow2(n)
rs = np.random.RandomState(4242)
t0 = time.time()
pts = be.points_from_uniform(rs.randint(0, 256, size=(2 * ng + 2, 64), dtype=np.uint8).tobytes())
enc = be.compress_points(pts)
cir = G.Circuit(be, n, Q, m, WL, WR, WO, WV, c)
gens = G.Generators(be, enc[:32], enc[32:64], [enc[64 + 32 * i: 96 + 32 * i] for i in range(ng)],
                    [enc[64 + 32 * (ng + i): 96 + 32 * (ng + i)] for i in range(ng)], cb)
be.synchronize()
setup_s = time.time() - t0
sb = lambda vals: b"".join(int(x).to_bytes(32, "little") for x in vals)
rows = []
for B in batches:
    aL, aR, aO, vv = [], [], [], []
    for i in range(B):
        perm = rs.permutation(k)
        x = int.from_bytes(rs.bytes(31), "little")
        v, a_L, a_R, a_O = W.shuffle_witness(k, perm, x)
        aL.append(sb(a_L)); aR.append(sb(a_R)); aO.append(sb(a_O)); vv.append(sb(v))
    gam = rs.randint(0, 256, size=(B * m, 32), dtype=np.uint8)
    gam[:, 31] &= 0x0F
    seeds = rs.randint(0, 256, size=(B, 32), dtype=np.uint8).tobytes()
    batch = G.Batch(be, cir, gens, B, "fixed", b"test")
    batch.upload_witness(b"".join(aL), b"".join(aR), b"".join(aO), gam.tobytes(), seeds)
    batch.commit(b"".join(vv))
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
    row = {"k": k, "n": n, "generators": 2 * ng + 2, "ipa_rounds": ng.bit_length() - 1, "batch": B, "table_window_bits": cb,
           "prove_ms": tp * 1e3, "verify_ms": tv * 1e3, "proofs_per_s": B / (tp + tv), "proof_bytes": batch.proof_len,
           "accepted": ok, "setup_s": setup_s}
    print(json.dumps(row), flush=True)
    rows.append(row)
    batch.free()
json.dump(rows, open("gpurun_out/large_deck.json", "w"), indent=1)
