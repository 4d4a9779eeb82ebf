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
