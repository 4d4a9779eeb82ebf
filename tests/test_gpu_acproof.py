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
    Vc = b"".join(R.compress(p) for p in V)
    proofs = G.prove_batch(backend, cir, gens, _sb(prover["a_L"]) * B, _sb(prover["a_R"]) * B, _sb(prover["a_O"]) * B,
                           _sb(prover["gamma"]) * B, b"".join(seeds), B, mode, V=Vc * B)
    plen = G.proof_len(n)
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
    Vc = b"".join(R.compress(p) for p in V)
    proofs = G.prove_batch(backend, cir, gens, _sb(prover["a_L"]) * B, _sb(prover["a_R"]) * B, _sb(prover["a_O"]) * B,
                           _sb(prover["gamma"]) * B, b"".join(seeds), B, "reference-fixed", V=Vc * B)
    plen = G.proof_len(104)
    for i, sd in enumerate(seeds):
        pb, ok, _, _ = _oracle(core, prover, V, sd, "reference-fixed", fast_msm)
        assert ok
        assert proofs[i * plen:(i + 1) * plen] == pb
    assert list(G.verify_batch(backend, cir, gens, proofs, Vc * B, B)) == [1, 1]


def test_tampered_proofs_and_commitments_are_rejected(backend):
    from bpperm_b200 import acproof as G
    core, prover, V, cir, gens = _setup(backend, 4, 9)
    n, m = core["n"], core["m"]
    B = 12
    seeds = b"".join(bytes([100 + i]) * 32 for i in range(B))
    Vc = bytearray(b"".join(R.compress(p) for p in V) * B)
    proofs = bytearray(G.prove_batch(backend, cir, gens, _sb(prover["a_L"]) * B, _sb(prover["a_R"]) * B,
                                     _sb(prover["a_O"]) * B, _sb(prover["gamma"]) * B, seeds, B, V=bytes(Vc)))
    plen = G.proof_len(n)
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
                       seeds[:32], 1, V=bytes(Vc[:32 * m]))
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
                           bytes([69]) * 32, 1, V=want)
        assert p2 == pb
        g2.free()


@pytest.mark.parametrize("k,B", [(2, 3), (5, 3), (52, 3), (300, 3), (70, 1030)])
def test_device_shuffle_witness_and_library_circuit_equal_the_oracle(backend, k, B):
    """SURVEY 8 row f-1 behind the C ABI: bpp_circuit_create_shuffle + bpp_acp_batch_gen_shuffle_witness against
    oracle/acproof.py shuffle_circuit / shuffle_witness - same commitments and byte-identical proofs as the uploaded
    witness over the Python-built circuit."""
    import numpy as np
    from bpperm_b200 import acproof as G
    from oracle import cref
    n, Q, m, WL, WR, WO, WV, c = A.shuffle_circuit(k)
    rs = np.random.RandomState(900 + k)
    enc = cref.compress(cref.from_uniform(rs.randint(0, 256, size=(2 * n + 2, 64), dtype=np.uint8).tobytes()))
    pts = [enc[32 * i:32 * i + 32] for i in range(2 * n + 2)]
    cir_py = G.Circuit(backend, n, Q, m, WL, WR, WO, WV, c)
    cir_lib = G.Circuit.shuffle(backend, k)
    assert (cir_lib.n, cir_lib.Q, cir_lib.m) == (n, Q, m)
    gens = G.Generators(backend, pts[0], pts[1], pts[2:2 + n], pts[2 + n:], 6)
    # B = 3: short chains go thread-per-chain (k <= 64), k = 300 through the block scan; (70, 1030): the thread-per-chain
    # form for medium chains once there are enough of them (k <= 1024 and 2 B >= 2048)
    deck = [(7 * i + 3) % L for i in range(1, k + 1)]
    perms, xs, wit = [], [], []
    for p in range(B):
        perm = rs.permutation(k)
        x = int.from_bytes(rs.bytes(32), "little") % L
        v = deck + [deck[j] for j in perm] + [x]
        aL, aR, aO = [0] * n, [0] * n, [0] * n
        for gb, vb in ((0, 0), (k - 1, k)):               # oracle.acproof.shuffle_witness, with this deck and permutation
            for i in range(k - 1):
                aL[gb + i] = (v[vb] - x) % L if i == 0 else aO[gb + i - 1]
                aR[gb + i] = (v[vb + i + 1] - x) % L
                aO[gb + i] = aL[gb + i] * aR[gb + i] % L
        perms.append(perm.astype(np.uint32).tobytes()); xs.append(x); wit.append((v, aL, aR, aO))
    gamma = [[int.from_bytes(rs.bytes(32), "little") % L for _ in range(m)] for _ in range(B)]
    seeds = b"".join((p + 1).to_bytes(4, "little") * 8 for p in range(B))
    gam_b = b"".join(_sb(g) for g in gamma)
    # uploaded witness, Python-built circuit
    b1 = G.Batch(backend, cir_py, gens, B)
    b1.upload_witness(b"".join(_sb(w[1]) for w in wit), b"".join(_sb(w[2]) for w in wit), b"".join(_sb(w[3]) for w in wit), gam_b, seeds)
    V1 = b1.commit(b"".join(_sb(w[0]) for w in wit))
    b1.prove()
    p1 = b1.download_proofs()
    # generated witness, library-built circuit
    b2 = G.Batch(backend, cir_lib, gens, B)
    b2.gen_shuffle_witness(_sb(deck), b"".join(perms), _sb(xs), gam_b, seeds)
    V2 = b2.commit(None)
    b2.prove()
    p2 = b2.download_proofs()
    assert V1 == V2 and p1 == p2
    b2.upload_proofs(p2, V2)
    b2.verify()
    assert b2.download_accept() == b"\x01" * B
    for o in (b1, b2, gens, cir_py, cir_lib):
        o.free()
