"""Proof wire format: the C ABI's to_wire / from_wire against the Python restatement of
bulletproofs 4.0.0 R1CSProof::to_bytes / from_bytes.  Byte handling only: runs without a GPU."""
import random

import pytest

from oracle import ristretto255 as R
from oracle import wire as W

import bpperm_b200
from bpperm_b200 import acproof as G
from bpperm_b200._lib import BppError

MODES = {0: "reference", 1: "reference-fixed", 2: "fixed"}


def _random_proof(rng, n, mode, canonical=True):
    plen = W.proof_len(n, mode)
    words = plen // 32
    out = bytearray()
    for i in range(words):
        if i < 8 or (mode == 2 and 11 <= i < words - 2):
            out += rng.randbytes(32)                      # points: any bytes, never looked at by the parser
        else:
            out += rng.randrange(R.L).to_bytes(32, "little")
    return bytes(out)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 8, 104, 200])
def test_round_trip_and_lengths(mode, n):
    rng = random.Random(1000 * mode + n)
    count = 5
    proofs = b"".join(_random_proof(rng, n, mode) for _ in range(count))
    assert G.proof_len(n, MODES[mode]) == W.proof_len(n, mode)
    assert G.wire_len(n, MODES[mode]) == W.proof_len(n, mode) + 1
    wire = G.to_wire(proofs, n, count, MODES[mode])
    plen = W.proof_len(n, mode)
    assert wire == b"".join(W.to_bytes(proofs[i * plen:(i + 1) * plen], mode) for i in range(count))
    back, status = G.from_wire(wire, n, count, MODES[mode])
    assert back == proofs and status == bytes(count)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_format_errors_match_the_restatement(mode):
    """Every single-field corruption: a wrong version byte, each scalar field set to l, l + 1, 2^256 - 1 and to l - 1
    (still canonical), and point fields set to non-canonical bytes (not a format error)."""
    n = 8
    rng = random.Random(77 + mode)
    good = _random_proof(rng, n, mode)
    plen, words = len(good), len(good) // 32
    cases = [W.to_bytes(good, mode)]
    for v in (0x00, 0x01, 0x80, 0x81, 0xFF):
        cases.append(bytes([v]) + good)
    for i in range(words):
        for val in (R.L, R.L + 1, 2 ** 256 - 1, R.L - 1, 2 ** 255 - 19):
            p = bytearray(good)
            p[32 * i:32 * i + 32] = val.to_bytes(32, "little")
            cases.append(W.to_bytes(bytes(p), mode))
    wire = b"".join(cases)
    back, status = G.from_wire(wire, n, len(cases), MODES[mode])
    n_bad = 0
    for k, rec in enumerate(cases):
        want = W.from_bytes(rec, n, mode)
        assert status[k] == (0 if want is not None else 1), k
        assert back[k * plen:(k + 1) * plen] == (want if want is not None else bytes(plen)), k
        n_bad += want is None
    assert 0 < n_bad < len(cases)


def test_wrong_record_length_and_arguments():
    n, mode = 8, 2
    rng = random.Random(5)
    rec = W.to_bytes(_random_proof(rng, n, mode), mode)
    for bad in (rec[:-1], rec + b"\x00", rec[:1], rec + bytes(32)):
        assert W.from_bytes(bad, n, mode) is None
        with pytest.raises(BppError):
            G.from_wire(bad, n, 1, "fixed")
    with pytest.raises(ValueError):
        G.to_wire(rec, n, 1, "fixed")          # not a bare proof
    # a record of another circuit size is a length mismatch, as in from_bytes' element-count check
    other = W.to_bytes(_random_proof(rng, 16, mode), mode)
    with pytest.raises(BppError):
        G.from_wire(other, n, 1, "fixed")
