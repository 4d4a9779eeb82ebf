"""GPU parity: the batched shuffle prover / verifier vs the oracle's restatement of circuit_lib.rs.
Proof bytes must be identical under the same seeded RNG and transcript label; accept/reject
decisions must be identical."""
import pytest

from oracle import acproof as A, ristretto255 as R
from oracle.chacha import ChaChaRng

pytestmark = pytest.mark.gpu
L = R.L


def _sb(v):
    return b"".join(R.sc_bytes(s) for s in v)


def _setup(backend, k, seed):
    from bpperm_b200 import acproof as G
    rng = ChaChaRng(bytes([seed]) * 32)
    core, prover, V = A.make_instance(k, rng)
    WL, WR, WO, WV = core["sparse"]
    cir = G.Circuit(backend, core["n"], core["Q"], core["m"], WL, WR, WO, WV, core["c_vec"])
    gens = G.Generators(backend, R.compress(core["g_base"]), R.compress(core["h_base"]),
                        [R.compress(p) for p in core["G_vec"]], [R.compress(p) for p in core["H_vec"]])
    return core, prover, V, cir, gens


def _oracle(core, prover, V, seed, mode, msm=None):
    return A.run_flow(core, prover, V, ChaChaRng(seed), mode, b"test", msm)


@pytest.mark.parametrize("k", [2, 3, 5])
@pytest.mark.parametrize("mode", ["reference-fixed", "reference"])
def test_small_decks_proof_bytes_and_decisions(backend, k, mode):
    from bpperm_b200 import acproof as G
    core, prover, V, cir, gens = _setup(backend, k, 40 + k)
    seeds = [bytes([i + 1]) * 32 for i in range(3)]
    B = len(seeds)
    n, m = core["n"], core["m"]
    proofs = G.prove_batch(backend, cir, gens, _sb(prover["a_L"]) * B, _sb(prover["a_R"]) * B, _sb(prover["a_O"]) * B,
                           _sb(prover["gamma"]) * B, b"".join(seeds), B, mode)
    plen = G.proof_len(n)
    Vc = b"".join(R.compress(p) for p in V)
    want_ok = []
    for i, sd in enumerate(seeds):
        pb, ok, _, _ = _oracle(core, prover, V, sd, mode)
        assert proofs[i * plen:(i + 1) * plen] == pb, (k, mode, i)
        want_ok.append(1 if ok else 0)
    acc = G.verify_batch(backend, cir, gens, proofs, Vc * B, B, mode)
    assert list(acc) == want_ok
    assert all(want_ok) == (mode == "reference-fixed")   # the reference's verifier never accepts


def test_52_card_shuffle_proof_is_byte_identical(backend):
    """BASELINE configs[1]: k = 52 -> n = 104, Q = 208, m = 105."""
    from bpperm_b200 import acproof as G
    from oracle import cref
    core, prover, V, cir, gens = _setup(backend, 52, 52)
    assert (core["n"], core["Q"], core["m"]) == (104, 208, 105)

    def fast_msm(scalars, points):   # the C restatement speeds the oracle's MSMs up; results identical
        enc = b"".join(R.compress(p) for p in points)
        out = cref.msm(b"".join(R.sc_bytes(s) for s in scalars), cref.decompress(enc))
        return R.decompress(out) if out != bytes(32) else R.IDENTITY

    seeds = [b"\x77" * 32, b"\x78" * 32]
    B = len(seeds)
    proofs = G.prove_batch(backend, cir, gens, _sb(prover["a_L"]) * B, _sb(prover["a_R"]) * B, _sb(prover["a_O"]) * B,
                           _sb(prover["gamma"]) * B, b"".join(seeds), B, "reference-fixed")
    plen = G.proof_len(104)
    for i, sd in enumerate(seeds):
        pb, ok, _, _ = _oracle(core, prover, V, sd, "reference-fixed", fast_msm)
        assert ok
        assert proofs[i * plen:(i + 1) * plen] == pb
    Vc = b"".join(R.compress(p) for p in V)
    assert list(G.verify_batch(backend, cir, gens, proofs, Vc * B, B)) == [1, 1]


def test_tampered_proofs_and_commitments_are_rejected(backend):
    from bpperm_b200 import acproof as G
    core, prover, V, cir, gens = _setup(backend, 4, 9)
    n, m = core["n"], core["m"]
    B = 12
    seeds = b"".join(bytes([100 + i]) * 32 for i in range(B))
    proofs = bytearray(G.prove_batch(backend, cir, gens, _sb(prover["a_L"]) * B, _sb(prover["a_R"]) * B,
                                     _sb(prover["a_O"]) * B, _sb(prover["gamma"]) * B, seeds, B))
    plen = G.proof_len(n)
    Vc = bytearray(b"".join(R.compress(p) for p in V) * B)
    good = bytes(proofs)
    assert list(G.verify_batch(backend, cir, gens, good, bytes(Vc), B)) == [1] * B
    # one corruption per proof, each in a different field
    fields = [0, 1, 2, 3, 7, 8, 9, 10, 11, 11 + n]          # A_I, A_O, S, T1, T6, tau_x, mu, t, l[0], r[0]
    for i, f in enumerate(fields):
        if f < 8:   # replace the point with a different valid point
            proofs[i * plen + 32 * f: i * plen + 32 * f + 32] = R.compress(R.pt_mul(i + 2, R.BASEPOINT))
        else:
            s = (int.from_bytes(proofs[i * plen + 32 * f: i * plen + 32 * f + 32], "little") + 1) % L
            proofs[i * plen + 32 * f: i * plen + 32 * f + 32] = R.sc_bytes(s)
    # proof 10: an invalid point encoding (dalek decompress -> None); proof 11: a wrong commitment V_0
    proofs[10 * plen + 96: 10 * plen + 128] = b"\x01" + bytes(31)
    Vc[11 * 32 * m: 11 * 32 * m + 32] = R.compress(R.pt_mul(5, R.BASEPOINT))
    acc = G.verify_batch(backend, cir, gens, bytes(proofs), bytes(Vc), B)
    assert list(acc) == [0] * B
    # wrong witness (a_O[0] + 1): the prover still produces bytes, the verifier must reject
    bad_aO = list(prover["a_O"])
    bad_aO[0] = (bad_aO[0] + 1) % L
    p2 = G.prove_batch(backend, cir, gens, _sb(prover["a_L"]), _sb(prover["a_R"]), _sb(bad_aO), _sb(prover["gamma"]),
                       seeds[:32], 1)
    _, ok, _, _ = _oracle(core, dict(prover, a_O=bad_aO), V, seeds[:32], "reference-fixed")
    assert not ok and list(G.verify_batch(backend, cir, gens, p2, bytes(Vc[:32 * m]), 1)) == [0]


def test_commit_variables_and_staged_batch(backend):
    from bpperm_b200 import acproof as G
    core, prover, V, cir, gens = _setup(backend, 3, 11)
    B = 70   # > 64 exercises the threaded transcript path
    batch = G.Batch(backend, cir, gens, B)
    seeds = b"".join(bytes([i]) * 32 for i in range(B))
    batch.upload_witness(_sb(prover["a_L"]) * B, _sb(prover["a_R"]) * B, _sb(prover["a_O"]) * B, _sb(prover["gamma"]) * B, seeds)
    Vc = batch.commit(_sb(prover["v"]) * B)
    want = b"".join(R.compress(p) for p in V)     # weights.rs:58-61 commit_variables
    assert Vc == want * B
    batch.prove()
    proofs = batch.download_proofs()
    batch.verify(b"\x09" * 32)
    assert batch.download_accept() == b"\x01" * B
    pb, ok, _, _ = _oracle(core, prover, V, bytes([69]) * 32, "reference-fixed")
    plen = G.proof_len(core["n"])
    assert ok and proofs[69 * plen:] == pb
    # all window widths of the fixed-base tables give the same bytes
    for c in (4, 5, 12):
        g2 = G.Generators(backend, R.compress(core["g_base"]), R.compress(core["h_base"]),
                          [R.compress(p) for p in core["G_vec"]], [R.compress(p) for p in core["H_vec"]], c)
        p2 = G.prove_batch(backend, cir, g2, _sb(prover["a_L"]), _sb(prover["a_R"]), _sb(prover["a_O"]), _sb(prover["gamma"]),
                           bytes([69]) * 32, 1)
        assert p2 == pb
        g2.free()
