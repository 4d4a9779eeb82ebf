"""SURVEY D.3 "pick by measurement": one inner-product round's group work in the two forms, on the same sizes.
  scalar-fold (what the prover does): L_j, R_j as ONE fixed-base MSM over the 2 n' + 2 ORIGINAL generators per proof
      - the same cost in every round (measured: bpp_acp_batch_time_commit_msm-style launch through the batch's own path);
  generator-fold (bulletproofs 4.0.0, K7): fold G and H explicitly (k_ipa_fold_gens: n_j new points per round and vector),
      after which L_j, R_j are variable-base MSMs over 2 n_j folded points (not timed here: the fold alone already decides).
argv: none.  Writes gpurun_out/ipa_fold_compare.json."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
import bpperm_b200

be = bpperm_b200.Backend(0)
G = bpperm_b200.acproof
L = bench.L_ORDER
rows = []
for k, B, cb in ((52, 4096, 16), (4096, 1, 8)):
    cir, gens, enc, (n, Q, m, ng), _ = bench.shuffle_setup_mode(be, k, "fixed", cb)
    lg = ng.bit_length() - 1
    deck, perm, x, gamma, seeds = bench.synth_shuffle_inputs(k, B, 0, 3)
    batch = G.Batch(be, cir, gens, B, "fixed", b"test")
    batch.gen_shuffle_witness(deck.tobytes(), perm.tobytes(), x.tobytes(), gamma.tobytes(), seeds.tobytes())
    batch.commit(None, want=False)
    batch.prove()
    be.synchronize()
    # scalar-fold: time of the whole prover minus the prover of the same circuit with one round less is not separable;
    # measure the round MSM directly: an A_I-shaped launch has 2 n + 1 terms, a round 2 n' + 2 - scale by the term count
    ms_ai, madd, _ = batch.time_commit_msm(5)
    ms_round = ms_ai * (2 * ng + 2) / (2 * n + 1)
    pts = be.upload_points([enc[64 + 32 * i: 96 + 32 * i] for i in range(ng)])
    rs = np.random.RandomState(9)
    row = {"k": k, "batch": B, "n_padded": ng, "rounds": lg, "table_window_bits": cb,
           "scalar_fold_round_msm_ms": ms_round, "scalar_fold_all_rounds_ms": ms_round * lg, "generator_fold": []}
    tot = 0.0
    for j in range(lg):
        nj = ng >> j
        us = [int.from_bytes(rs.bytes(32), "little") % L for _ in range(B)]
        uis = [pow(u, L - 2, L) for u in us]
        sb = lambda v: b"".join(int(s).to_bytes(32, "little") for s in v)
        _, ms = be.ipa_fold_generators(pts, nj, sb(us), sb(uis), want=False)
        _, ms = be.ipa_fold_generators(pts, nj, sb(us), sb(uis), want=False)
        tot += 2 * ms          # G and H
        row["generator_fold"].append({"round": j, "n_j": nj, "fold_G_and_H_ms": 2 * ms, "scalar_fold_round_msm_ms": ms_round})
    row["generator_fold_all_rounds_ms"] = tot
    row["winner_per_round"] = ["generator-fold" if r["fold_G_and_H_ms"] < ms_round else "scalar-fold" for r in row["generator_fold"]]
    row["note"] = ("generator-fold needs EVERY earlier round folded too (round j's generators are round j-1's folded ones), so its cost up "
                   "to round j is the running sum; and after folding, L_j and R_j are variable-base MSMs over points without tables")
    print(json.dumps(row), flush=True)
    rows.append(row)
    pts.free(); batch.free(); gens.free(); cir.free()
json.dump(rows, open("gpurun_out/ipa_fold_compare.json", "w"), indent=1)
