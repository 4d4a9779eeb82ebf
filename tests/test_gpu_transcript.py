"""GPU parity: Merlin transcripts on the device (merlin_dev.cuh) against the published Merlin KAT, the
oracle's STROBE-128 restatement (oracle/merlin.py) and the library's host transcripts."""
import random

import pytest

from oracle import acproof as A, merlin as M, ristretto255 as R
from oracle.chacha import ChaChaRng

pytestmark = pytest.mark.gpu


def _run(backend, records):
    from bpperm_b200 import acproof as G
    return G.transcript_script(backend, records)


def _oracle(records):
    t = M.Transcript(records[0][2])
    out = b""
    for op, label, arg in records[1:]:
        if op == "append":
            t.append_message(label, arg)
        else:
            out += t.challenge_bytes(label, arg)
    return out


def test_merlin_published_kat(backend):
    """merlin 3.0.0 tests::equivalence_simple"""
    rec = [("append", b"dom-sep", b"test protocol"), ("append", b"some label", b"some data"), ("challenge", b"challenge", 32)]
    got = _run(backend, rec)
    assert got.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    assert got == _oracle(rec)


def test_random_scripts_match_oracle(backend):
    """message and challenge lengths straddling the 166-byte STROBE rate, empty messages, many operations"""
    rnd = random.Random(5)
    for case in range(12):
        rec = [("append", b"dom-sep", bytes(rnd.getrandbits(8) for _ in range(rnd.choice([0, 4, 13, 200]))))]
        for _ in range(rnd.randint(3, 25)):
            label = bytes(rnd.getrandbits(8) for _ in range(rnd.choice([1, 3, 12])))
            if rnd.random() < 0.6:
                n = rnd.choice([0, 1, 32, 64, 159, 160, 161, 165, 166, 167, 331, 332, 333, 500])
                rec.append(("append", label, bytes(rnd.getrandbits(8) for _ in range(n))))
            else:
                rec.append(("challenge", label, rnd.choice([1, 32, 64, 165, 166, 167, 400])))
        rec.append(("challenge", b"last", 64))
        assert _run(backend, rec) == _oracle(rec), case


@pytest.mark.parametrize("mode", ["reference-fixed", "reference", "fixed"])
def test_device_and_host_transcripts_agree(backend, mode):
    """the same batch proved and verified with Fiat-Shamir on the device and on host threads: identical proof
    bytes and decisions (and, through tests/test_gpu_acproof.py / test_gpu_ipa.py, identical to the oracle's)"""
    from bpperm_b200 import acproof as G
    k = 5
    rng = ChaChaRng(bytes([77]) * 32)
    if mode == "fixed":
        from oracle import ipa
        core, prover, V = ipa.make_instance(k, rng, dense_weights=True)   # next_pow2(n) generators
    else:
        core, prover, V = A.make_instance(k, rng)
    n = core["n"]
    WL, WR, WO, WV = core["sparse"]
    cir = G.Circuit(backend, n, core["Q"], core["m"], WL, WR, WO, WV, core["c_vec"])
    gens = G.Generators(backend, R.compress(core["g_base"]), R.compress(core["h_base"]),
                        [R.compress(p) for p in core["G_vec"]], [R.compress(p) for p in core["H_vec"]])
    B = 37
    sb = lambda v: b"".join(R.sc_bytes(s) for s in v)
    seeds = b"".join(bytes([i + 1]) * 32 for i in range(B))
    Vc = b"".join(R.compress(p) for p in V) * B
    res = []
    for host in (False, True):
        batch = G.Batch(backend, cir, gens, B, mode, b"test")
        batch.set_host_transcripts(host)
        batch.upload_witness(sb(prover["a_L"]) * B, sb(prover["a_R"]) * B, sb(prover["a_O"]) * B, sb(prover["gamma"]) * B, seeds)
        batch.upload_commitments(Vc)
        batch.prove()
        proofs = batch.download_proofs()
        bad = bytearray(proofs)
        bad[batch.proof_len * 3 + 40] ^= 1       # corrupt A_O of proof 3
        batch.upload_proofs(bytes(bad), Vc)
        batch.verify(b"\x11" * 32)
        res.append((proofs, batch.download_accept()))
        batch.free()
    assert res[0][0] == res[1][0]
    assert res[0][1] == res[1][1]
    want = [0 if (i == 3 or mode == "reference") else 1 for i in range(B)]
    assert list(res[0][1]) == want


@pytest.mark.parametrize("mode", ["reference-fixed", "fixed"])
def test_batch_rlc_verification_matches_per_proof_decisions(backend, mode):
    """one random-linear-combination MSM over the batch (default) vs per-proof verification: same accept bytes
    for an all-valid batch (combined check decides) and for batches with corrupted proofs (falls back)"""
    from bpperm_b200 import acproof as G
    k = 4
    rng = ChaChaRng(bytes([91]) * 32)
    if mode == "fixed":
        from oracle import ipa
        core, prover, V = ipa.make_instance(k, rng, dense_weights=True)
    else:
        core, prover, V = A.make_instance(k, rng)
    n, m = core["n"], core["m"]
    WL, WR, WO, WV = core["sparse"]
    cir = G.Circuit(backend, n, core["Q"], m, WL, WR, WO, WV, core["c_vec"])
    gens = G.Generators(backend, R.compress(core["g_base"]), R.compress(core["h_base"]),
                        [R.compress(p) for p in core["G_vec"]], [R.compress(p) for p in core["H_vec"]])
    B = 64      # 64 x (m + 8) >= 1024 points: the combined check is active
    sb = lambda v: b"".join(R.sc_bytes(s) for s in v)
    seeds = b"".join(bytes([i + 1]) * 32 for i in range(B))
    Vc = b"".join(R.compress(p) for p in V) * B
    batch = G.Batch(backend, cir, gens, B, mode, b"test")
    batch.upload_witness(sb(prover["a_L"]) * B, sb(prover["a_R"]) * B, sb(prover["a_O"]) * B, sb(prover["gamma"]) * B, seeds)
    batch.upload_commitments(Vc)
    batch.prove()
    proofs = batch.download_proofs()
    plen = batch.proof_len
    other = R.compress(R.pt_double(R.BASEPOINT))

    def decisions(pr, Vs, rlc):
        batch.set_batch_rlc(rlc)
        batch.upload_proofs(pr, Vs)
        l0 = backend.launch_count
        batch.verify(b"\x33" * 32)
        return batch.download_accept(), backend.launch_count - l0

    ok_rlc, n_rlc = decisions(proofs, Vc, True)
    ok_pp, n_pp = decisions(proofs, Vc, False)
    assert ok_rlc == ok_pp == b"\x01" * B
    cases = []
    for victim, off in ((0, 0), (5, 3 * 32), (B - 1, 8 * 32 + 5), (7, plen - 1)):    # A_I, T_1, a proof scalar, last byte
        bad = bytearray(proofs)
        if off % 32 == 0 and off < 256:
            bad[victim * plen + off:victim * plen + off + 32] = other     # another valid point
        else:
            bad[victim * plen + off] ^= 1
        cases.append((bytes(bad), Vc, victim, mode == "fixed" or off != plen - 1))
    badV = bytearray(Vc)
    badV[32 * (m * 9 + 2):32 * (m * 9 + 3)] = other                        # commitment 2 of proof 9
    cases.append((proofs, bytes(badV), 9, True))
    for pr, Vs, victim, needs_msm in cases:
        a1, n1 = decisions(pr, Vs, True)
        a0, _ = decisions(pr, Vs, False)
        assert a1 == a0
        assert list(a1) == [0 if i == victim else 1 for i in range(B)], victim
        if needs_msm:   # (a flipped byte of r already fails t == <l, r>: weight 0, no fall-back needed)
            assert n1 > n_rlc     # the combined check failed and the per-proof kernels ran
    batch.free()


def test_offsets_that_cancelled_under_the_round_1_weights_are_rejected(backend):
    """the round-1 verifier drew its per-proof weight w_p and batch weight rho_p from
    ChaCha20(verifier_seed, block p) alone.  Whoever knew the seed could give two proofs of one batch tau_x offsets
    d_0, d_1 with rho_0 w_0 d_0 + rho_1 w_1 d_1 = 0 - the h-coefficient of the combined check is sum_p rho_p (w_p tau_x_p
    - mu_p) - and both were accepted.  The weights are now challenge scalars of each proof's own transcript (which has
    absorbed V and every proof byte) rekeyed with the seed: the same crafted pair, under the same publicly known seed,
    must be rejected, with the combined check and per proof."""
    from bpperm_b200 import acproof as G
    from oracle.chacha import chacha20_block
    k = 4
    core, prover, V = A.make_instance(k, ChaChaRng(bytes([123]) * 32))
    n, m = core["n"], core["m"]
    WL, WR, WO, WV = core["sparse"]
    cir = G.Circuit(backend, n, core["Q"], m, WL, WR, WO, WV, core["c_vec"])
    gens = G.Generators(backend, R.compress(core["g_base"]), R.compress(core["h_base"]),
                        [R.compress(p) for p in core["G_vec"]], [R.compress(p) for p in core["H_vec"]])
    B = 64
    sb = lambda v: b"".join(R.sc_bytes(s) for s in v)
    seeds = b"".join(bytes([i + 1]) * 32 for i in range(B))
    Vc = b"".join(R.compress(p) for p in V) * B
    batch = G.Batch(backend, cir, gens, B, "reference-fixed", b"test")
    batch.upload_witness(sb(prover["a_L"]) * B, sb(prover["a_R"]) * B, sb(prover["a_O"]) * B, sb(prover["gamma"]) * B, seeds)
    batch.upload_commitments(Vc)
    batch.prove()
    proofs = bytearray(batch.download_proofs())
    plen = batch.proof_len
    seed = b"\x33" * 32                      # known to the attacker
    old_w = [R.sc_from_wide(chacha20_block(seed, p, 0)) for p in range(2)]      # round-1 k_acp_weights: stream word 0
    old_rho = [R.sc_from_wide(chacha20_block(seed, p, 1)) for p in range(2)]    # round-1 k_rlc_weights: stream word 1
    d0 = 0x1234567
    d1 = (-old_rho[0] * old_w[0] * d0) * R.sc_inv(old_rho[1] * old_w[1] % R.L) % R.L
    for p, d in ((0, d0), (1, d1)):
        off = p * plen + 32 * 8             # tau_x: field 8 of a mode-1 proof
        t = (int.from_bytes(proofs[off:off + 32], "little") + d) % R.L
        proofs[off:off + 32] = R.sc_bytes(t)
    for rlc in (True, False):
        batch.set_batch_rlc(rlc)
        batch.upload_proofs(bytes(proofs), Vc)
        batch.verify(seed)
        acc = batch.download_accept()
        assert list(acc[:2]) == [0, 0] and acc[2:] == b"\x01" * (B - 2), rlc
    batch.free()
