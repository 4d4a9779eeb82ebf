"""CPU tests: pin the oracle (Python and C restatements) against published vectors and the
committed libsodium fixtures.  The reference itself holds no golden vectors."""
import hashlib
import json
import os

import pytest

from oracle import cref, merlin, ristretto255 as R
from oracle.chacha import ChaChaRng, chacha20_block

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "libsodium_ristretto255.json")))

RFC9496_MULTIPLES = [
    "0000000000000000000000000000000000000000000000000000000000000000",
    "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",
    "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919",
    "94741f5d5d52755ece4f23f044ee27d5d1ea1e2bd196b462166b16152a9d0259",
    "da80862773358b466ffadfe0b3293ab3d9fd53c5ea6c955358f568322daf6a57",
    "e882b131016b52c1d3337080187cf768423efccbb517bb495ab812c4160ff44e",
    "f64746d3c92b13050ed8d80236a7f0007c3b3f962f5ba793d19a601ebb1df403",
    "44f53520926ec81fbd5a387845beb7df85a96a24ece18738bdcfa6a7822a176d",
    "903293d8f2287ebe10e2374dc1a53e0bc887e592699f02d077d5263cdd55601c",
    "02622ace8f7303a31cafc63f8fc48fdc16e1c8c8d234b2f0d6685282a9076031",
    "20706fd788b2720a1ed2a5dad4952b01f413bcf0e7564de8cdc816689e2db95f",
    "bce83f8ba5dd2fa572864c24ba1810f9522bc6004afe95877ac73241cafdab42",
    "e4549ee16b9aa03099ca208c67adafcafa4c3f3e4e5303de6026e3ca8ff84460",
    "aa52e000df2e16f55fb1032fc33bc42742dad6bd5a8fc0be0167436c5948501f",
    "46376b80f409b29dc2b5f6f0c52591990896e5716f41477cd30085ab7f10301e",
    "e0c418f7c8d9c4cdd7395b93ea124f3ad99021bb681dfc3302a9d99a2e53e64e",
]


def test_rfc9496_generator_multiples():
    acc = R.IDENTITY
    for k, want in enumerate(RFC9496_MULTIPLES):
        assert R.compress(acc).hex() == want, k
        assert R.compress(R.pt_mul(k, R.BASEPOINT)).hex() == want
        acc = R.pt_add(acc, R.BASEPOINT)
    # C restatement: decode/encode round trip of the same vectors
    enc = b"".join(bytes.fromhex(h) for h in RFC9496_MULTIPLES)
    assert cref.compress(cref.decompress(enc)) == enc


def test_rfc9496_one_way_map():
    h = hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest()
    want = "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"
    assert R.compress(R.from_uniform_bytes(h)).hex() == want
    assert cref.compress(cref.from_uniform(h)).hex() == want


def test_rfc9496_bad_encodings_rejected():
    from test_gpu_field_group import RFC9496_BAD_ENCODINGS
    for h in RFC9496_BAD_ENCODINGS:
        assert R.decompress(bytes.fromhex(h)) is None
        with pytest.raises(ValueError):
            cref.decompress(bytes.fromhex(h))


def test_merlin_published_kat():
    t = merlin.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_keccak_matches_hashlib_sha3():
    # one-block SHA3-256 through our permutation == hashlib (independent implementation)
    import struct
    msg = b"bulletproof-perm"
    st = bytearray(200)
    st[:len(msg)] = msg
    st[len(msg)] ^= 0x06
    st[135] ^= 0x80
    lanes = merlin.keccak_f1600(list(struct.unpack("<25Q", bytes(st))))
    assert struct.pack("<25Q", *lanes)[:32] == hashlib.sha3_256(msg).digest()


def test_chacha20_stream():
    # RFC 8439-style block function, zero key/nonce keystream
    assert chacha20_block(bytes(32), 0).hex().startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    rng = ChaChaRng(bytes(32))
    assert R.sc_bytes(rng.scalar()).hex() == "4a53c3fbbc59970ee5f85af813875dffc13a904a2e53ae7e65fa0dea6e62c901"
    try:
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms
    except Exception:
        return
    key = bytes(range(32))
    enc = Cipher(algorithms.ChaCha20(key, bytes(16)), mode=None).encryptor()
    assert enc.update(bytes(256)) == ChaChaRng(key).fill_bytes(256)


def test_survey_e2_e3_e4_vectors():
    rng = ChaChaRng(bytes(range(32)))
    pts = [rng.point() for _ in range(4)]
    sc = [rng.scalar() for _ in range(4)]
    assert R.compress(pts[1]).hex() == "cad1e2e19eb6e2d414c2f16ce23bcf8712adbbe813bbfaeed4d93f338df1205d"
    for f in (R.msm_naive, R.msm_straus, R.msm_pippenger, R.vartime_multiscalar_mul):
        assert R.compress(f(sc, pts)).hex() == "ac0188282e26885b30102aa5ee4e91734b3328dba689b17245af361585d58d6f"
    t = merlin.Transcript(b"test")
    t.arithmetic_domain_sep(104)
    for lab, k in ((b"A_I", 1), (b"A_O", 2), (b"S", 3)):
        t.append_point(lab, bytes.fromhex(RFC9496_MULTIPLES[k]))
    y = t.challenge_scalar(b"y")
    z = t.challenge_scalar(b"z")
    assert R.sc_bytes(y).hex() == "134ddbf9759905621d6baea3534e91e67a9ed02d27571d58f1dac13b401ad000"
    assert R.sc_bytes(z).hex() == "1494a5770e79396c6ab575b7dff3f3b4b9a34f04f4f03899988a5dd3d8642701"


def test_libsodium_fixtures_python_oracle():
    for v in GOLD["from_hash"]:
        assert R.compress(R.from_uniform_bytes(bytes.fromhex(v["in"]))).hex() == v["out"]
    for v in GOLD["scalar_reduce"]:
        assert R.sc_bytes(R.sc_from_wide(bytes.fromhex(v["in"]))).hex() == v["out"]
    for v in GOLD["scalarmult"]:
        p = R.decompress(bytes.fromhex(v["p"]))
        s = int.from_bytes(bytes.fromhex(v["s"]), "little")
        assert R.compress(R.pt_mul(s, p)).hex() == v["out"]
    for v in GOLD["add"]:
        assert R.compress(R.pt_add(R.decompress(bytes.fromhex(v["p"])), R.decompress(bytes.fromhex(v["q"])))).hex() == v["out"]
    for v in GOLD["scalar_ops"]:
        a = int.from_bytes(bytes.fromhex(v["a"]), "little")
        b = int.from_bytes(bytes.fromhex(v["b"]), "little")
        assert R.sc_bytes(a * b).hex() == v["mul"] and R.sc_bytes(a + b).hex() == v["add"]
        assert R.sc_bytes(R.sc_inv(a)).hex() == v["inv_a"]
    pts = [R.decompress(bytes.fromhex(v["out"])) for v in GOLD["from_hash"]]
    sc = [int.from_bytes(bytes.fromhex(v["out"]), "little") for v in GOLD["scalar_reduce"]]
    for v in GOLD["msm"]:
        n = v["n"]
        assert R.compress(R.vartime_multiscalar_mul(sc[:n], pts[:n])).hex() == v["out"]


def test_libsodium_fixtures_c_oracle():
    blobs = b"".join(bytes.fromhex(v["in"]) for v in GOLD["from_hash"])
    enc = b"".join(bytes.fromhex(v["out"]) for v in GOLD["from_hash"])
    pts = cref.from_uniform(blobs)
    assert cref.compress(pts) == enc
    assert cref.compress(cref.decompress(enc)) == enc
    sc = b"".join(bytes.fromhex(v["out"]) for v in GOLD["scalar_reduce"])
    for v in GOLD["msm"]:
        n = v["n"]
        for forced in (-1, 0, 1):
            assert cref.msm(sc[:32 * n], pts[:160 * n], forced=forced).hex() == v["out"]


def test_c_oracle_equals_python_oracle_on_dalek_dispatch_sizes():
    rng = ChaChaRng(b"\x07" * 32)
    n = 520
    blobs = [rng.fill_bytes(64) for _ in range(n)]
    sc = [rng.scalar() for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, 1, R.L - 1
    pts_c = cref.from_uniform(b"".join(blobs))
    scb = b"".join(R.sc_bytes(s) for s in sc)
    pts_py = [R.from_uniform_bytes(b) for b in blobs]
    # the reference's MSM sizes at 52 cards + both sides of dalek's 190 / 500 thresholds
    for m in (2, 104, 105, 110, 111, 189, 190, 209, 499, 500, 520):
        want = R.compress(R.msm_naive(sc[:m], pts_py[:m]))
        assert cref.msm(scb[:32 * m], pts_c[:160 * m]) == want
    # python restatements of Straus / Pippenger (structure-level copies of dalek's algorithms)
    assert R.compress(R.msm_straus(sc[:40], pts_py[:40])) == R.compress(R.msm_naive(sc[:40], pts_py[:40]))
    assert R.compress(R.msm_pippenger(sc[:200], pts_py[:200])) == R.compress(R.msm_naive(sc[:200], pts_py[:200]))


def test_radix_recodings_reconstruct():
    rng = ChaChaRng(b"\x09" * 32)
    for _ in range(50):
        s = rng.scalar()
        for w in range(4, 17):
            d = R.to_radix_2w(s, w)
            assert sum(di << (w * i) for i, di in enumerate(d)) == s
            assert all(-(1 << (w - 1)) <= di <= (1 << (w - 1)) for di in d)
        naf = R.non_adjacent_form(s, 5)
        assert sum(di << i for i, di in enumerate(naf)) == s
        assert all(di == 0 or (di & 1 and abs(di) < 16) for di in naf)


def test_ristretto_coset_invariance():
    # P, P + (4-torsion) compress identically (RFC 9496: the encoding is of the coset)
    rng = ChaChaRng(b"\x0a" * 32)
    t4 = (R.SQRT_M1, 0, 1, 0)          # a point of order 4
    t2 = (0, R.P - 1, 1, 0)            # the point of order 2
    for _ in range(10):
        p = rng.point()
        e = R.compress(p)
        assert R.compress(R.pt_add(p, t4)) == e and R.compress(R.pt_add(p, t2)) == e
        assert R.pt_eq(p, R.pt_add(p, t4))
