"""CPU tests for the `fixed` protocol mode (standard powers + inner-product argument): the Python restatement
(oracle/ipa.py) and the C restatement (oracle/c/acproof_ref.c) produce the same bytes and decisions, proofs are
complete and tampering is rejected.  The reference has no IPA: parity for this mode is
oracle == oracle/c == CUDA, unpinned."""
import pytest

from oracle import cref, ipa, ristretto255 as R
from oracle.chacha import ChaChaRng
from oracle.merlin import Transcript

L = R.L


def _sb(v):
    return b"".join(R.sc_bytes(s) for s in v)


def c_instance(core):
    return cref.AcpFixedInstance.from_core(core)


def test_inner_product_proof_alone_verifies():
    """InnerProductProof::create then the verification equation, n = 8, with non-trivial generator factors."""
    rng = ChaChaRng(b"\x21" * 32)
    n = 8
    G = [rng.point() for _ in range(n)]
    H = [rng.point() for _ in range(n)]
    Q = rng.point()
    a = [rng.scalar() for _ in range(n)]
    b = [rng.scalar() for _ in range(n)]
    y_inv = rng.scalar()
    gf, hf = [1] * n, ipa.std_powers(y_inv, n)
    P = R.vartime_multiscalar_mul([x * f % L for x, f in zip(a, gf)] + [x * f % L for x, f in zip(b, hf)] + [ipa.inner_product(a, b)],
                                  G + H + [Q])
    t1, t2 = Transcript(b"ipp test"), Transcript(b"ipp test")
    pr = ipa.InnerProductProof.create(t1, Q, gf, hf, G, H, a, b)
    assert len(pr.L_vec) == 3
    u_sq, u_inv_sq, s = pr.verification_scalars(n, t2)
    # s[i] * s[n-1-i] = 1 and s[0] = prod u_j^-1
    assert all(s[i] * s[n - 1 - i] % L == 1 for i in range(n))
    rhs = R.vartime_multiscalar_mul([pr.a * si % L for si in s] + [pr.b * s[n - 1 - i] % L * hf[i] % L for i in range(n)] +
                                    [pr.a * pr.b % L] + [(L - v) % L for v in u_sq] + [(L - v) % L for v in u_inv_sq],
                                    G + H + [Q] + [R.decompress(x) for x in pr.L_vec] + [R.decompress(x) for x in pr.R_vec])
    assert R.pt_eq(P, rhs)
    assert t1.challenge_bytes(b"x", 16) == t2.challenge_bytes(b"x", 16)    # prover and verifier transcripts agree


@pytest.mark.parametrize("k", [2, 3, 5, 9])
def test_fixed_mode_python_equals_c_and_is_complete_and_sound(k):
    rng = ChaChaRng(bytes([k]) * 32)
    core, prover, V = ipa.make_instance(k, rng)
    inst = c_instance(core)
    Vp = inst.commit(_sb(prover["v"]), _sb(prover["gamma"]))
    assert cref.compress(Vp) == b"".join(R.compress(p) for p in V)
    seed = bytes([k + 1]) * 32
    proof, st = ipa.prove(core, prover, V, ChaChaRng(seed))
    assert len(proof) == ipa.proof_len(core["n"]) == inst.proof_len
    assert inst.prove(_sb(prover["a_L"]), _sb(prover["a_R"]), _sb(prover["a_O"]), _sb(prover["gamma"]), Vp, seed) == proof
    assert ipa.verify(core, V, proof) and inst.verify(proof, Vp)
    assert st["t_hat"] == ipa.inner_product(st["l"], st["r"])
    npad = ipa.next_pow2(core["n"])
    assert st["l"][core["n"]:] == [0] * (npad - core["n"])
    words = len(proof) // 32
    for f in range(words):          # flip one bit in every 32-byte field: both verifiers reject
        bad = bytearray(proof)
        bad[32 * f + 3] ^= 0x10
        bad = bytes(bad)
        assert not inst.verify(bad, Vp), f
        if f % 3 == 0:
            assert not ipa.verify(core, V, bad), f
    # identity L_0 (validate_and_append_point) and a different label
    bad = proof[:32 * 11] + bytes(32) + proof[32 * 12:]
    assert not ipa.verify(core, V, bad) and not inst.verify(bad, Vp)
    assert not ipa.verify(core, V, proof, b"other") and not inst.verify(proof, Vp, b"other")


def test_fixed_mode_rejects_a_non_permutation():
    from oracle import acproof as A
    k = 4
    rng = ChaChaRng(b"\x44" * 32)
    core, prover, V = ipa.make_instance(k, rng)
    bad_v = list(prover["v"])
    bad_v[k] = (bad_v[k] + 1) % L
    x = bad_v[-1]
    a_L, a_R, a_O = [0] * (2 * k), [0] * (2 * k), [0] * (2 * k)
    for gb, vb in ((0, 0), (k - 1, k)):
        for i in range(k - 1):
            a_L[gb + i] = (bad_v[vb] - x) % L if i == 0 else a_O[gb + i - 1]
            a_R[gb + i] = (bad_v[vb + i + 1] - x) % L
            a_O[gb + i] = a_L[gb + i] * a_R[gb + i] % L
    V2 = A.commit_variables(bad_v, prover["gamma"], core["g_base"], core["h_base"])
    proof, _ = ipa.prove(core, dict(prover, a_L=a_L, a_R=a_R, a_O=a_O, v=bad_v), V2, ChaChaRng(b"\x01" * 32))
    assert not ipa.verify(core, V2, proof)
