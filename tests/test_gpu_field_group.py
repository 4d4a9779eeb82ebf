"""GPU parity: field / group / encoding primitives vs the CPU oracle, through the C ABI
(bpp_test_op, bpp_points_*).  Bit-exact: every comparison is on canonical 32-byte encodings."""
import random

import pytest

from oracle import ristretto255 as R
from oracle.chacha import ChaChaRng

pytestmark = pytest.mark.gpu

P = R.P
FE_MUL, FE_ADD, FE_SUB, FE_INVERT, FE_CANON, GE_ADD, GE_DOUBLE, GE_ROUNDTRIP, GE_SCALARMULT, FE_SQR, GE_DOUBLE_Z1 = range(11)


def _le(x):
    return x.to_bytes(32, "little")


def _edge_values():
    return [0, 1, 2, 19, 37, 38, P - 1, P, P + 1, P + 18, P + 19, 2**255 - 1, 2**255, 2**255 + 18, 2**255 + 19,
            2 * P, 2 * P + 1, 2 * P + 37, 2**256 - 1, 2**256 - 38, 2**256 - 39, 2**32 - 1, 2**32, 2**224,
            (2**256 - 1) ^ (2**32 - 1), 2**256 - 2**32, 0xFFFFFFDA, 0xFFFFFFDB]


@pytest.mark.parametrize("op,fn", [(FE_MUL, lambda a, b: a * b), (FE_ADD, lambda a, b: a + b),
                                   (FE_SUB, lambda a, b: a - b), (FE_CANON, lambda a, b: a)])
def test_field_ops_match_oracle(backend, op, fn):
    rnd = random.Random(1234 + op)
    edges = _edge_values()
    pairs = [(a, b) for a in edges for b in edges]
    pairs += [(rnd.getrandbits(256), rnd.getrandbits(256)) for _ in range(20000)]
    a = b"".join(_le(x) for x, _ in pairs)
    b = b"".join(_le(y) for _, y in pairs)
    out = backend.test_op(op, a, b)
    for i, (x, y) in enumerate(pairs):
        assert out[32 * i:32 * i + 32] == _le(fn(x, y) % P), (op, hex(x), hex(y))


def test_field_invert(backend):
    rnd = random.Random(99)
    xs = _edge_values() + [rnd.getrandbits(256) for _ in range(500)]
    a = b"".join(_le(x) for x in xs)
    out = backend.test_op(FE_INVERT, a, a)
    for i, x in enumerate(xs):
        assert out[32 * i:32 * i + 32] == _le(pow(x % P, P - 2, P))


def test_field_square(backend):
    """dedicated squaring (28 doubled off-diagonal products + 8 squares) against x*x, edge limbs included"""
    rnd = random.Random(1234)
    xs = _edge_values() + [rnd.getrandbits(256) for _ in range(2000)]
    for _ in range(1000):  # limbs drawn from {0, 1, 2^31, 2^32-2, 2^32-1, random}: worst cases for the carry chains
        xs.append(sum(rnd.choice([0, 1, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF, rnd.getrandbits(32)]) << (32 * i)
                      for i in range(8)))
    a = b"".join(_le(x) for x in xs)
    out = backend.test_op(FE_SQR, a, a)
    for i, x in enumerate(xs):
        assert out[32 * i:32 * i + 32] == _le(x * x % P), hex(x)


def _points(n, seed):
    rng = ChaChaRng(bytes([seed]) * 32)
    return [rng.point() for _ in range(n)]


def test_point_add_double_roundtrip(backend):
    pts = _points(200, 7) + [R.IDENTITY, R.BASEPOINT]
    qs = _points(200, 8) + [R.BASEPOINT, R.IDENTITY]
    a = b"".join(R.compress(p) for p in pts)
    b = b"".join(R.compress(q) for q in qs)
    out = backend.test_op(GE_ADD, a, b)
    out2 = backend.test_op(GE_DOUBLE, a, b)
    out3 = backend.test_op(GE_ROUNDTRIP, a, b)
    out4 = backend.test_op(GE_DOUBLE_Z1, a, b)
    for i, (p, q) in enumerate(zip(pts, qs)):
        assert out[32 * i:32 * i + 32] == R.compress(R.pt_add(p, q))
        assert out2[32 * i:32 * i + 32] == R.compress(R.pt_double(p))
        assert out3[32 * i:32 * i + 32] == R.compress(p)
        assert out4[32 * i:32 * i + 32] == R.compress(R.pt_double(p))


def test_point_plus_negative_is_identity(backend):
    pts = _points(20, 9)
    a = b"".join(R.compress(p) for p in pts)
    b = b"".join(R.compress(R.pt_neg(p)) for p in pts)
    assert backend.test_op(GE_ADD, a, b) == bytes(32 * len(pts))


def test_scalar_mult_matches_oracle(backend):
    rng = ChaChaRng(b"\x05" * 32)
    pts = [rng.point() for _ in range(64)]
    ks = [rng.scalar() for _ in range(60)] + [0, 1, R.L - 1, 2**252]
    out = backend.test_op(GE_SCALARMULT, b"".join(R.compress(p) for p in pts), b"".join(R.sc_bytes(k) for k in ks))
    for i, (p, k) in enumerate(zip(pts, ks)):
        assert out[32 * i:32 * i + 32] == R.compress(R.pt_mul(k, p))


def test_rfc9496_multiples_of_generator(backend):
    # RFC 9496 A.1: encodings of 0*B .. 15*B (first four checked literally, all against the oracle)
    want = ["0000000000000000000000000000000000000000000000000000000000000000",
            "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",
            "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919",
            "94741f5d5d52755ece4f23f044ee27d5d1ea1e2bd196b462166b16152a9d0259"]
    base = R.compress(R.BASEPOINT)
    out = backend.test_op(GE_SCALARMULT, base * 16, b"".join(_le(k) for k in range(16)))
    for k in range(16):
        enc = out[32 * k:32 * k + 32]
        assert enc == R.compress(R.pt_mul(k, R.BASEPOINT))
        if k < 4:
            assert enc.hex() == want[k]


RFC9496_BAD_ENCODINGS = [
    # non-canonical field encodings
    "00ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff",
    "ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "f3ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "edffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    # negative field elements
    "0100000000000000000000000000000000000000000000000000000000000000",
    "01ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "ed57ffd8c914fb201471d1c3d245ce3c746fcbe63a3679d51b6a516ebebe0e20",
    "c34c4e1826e5d403b78e246e88aa051c36ccf0aafebffe137d148a2bf9104562",
    # non-square x^2
    "26948d35ca62e643e26a83177332e6b6afeb9d08e4268b650f1f5bbd8d81d371",
    "4eac077a713c57b4f4397629a4145982c661f48044dd3f96427d40b147d9742f",
    # negative xy
    "3eb858e78f5a7254d8c9731174a94f76755fd3941c0ac93735c07ba14579630e",
    "a45fdc55c76448c049a1ab33f17023edfb2be3581e9c7aade8a6125215e04220",
    # s = -1, which causes y = 0
    "ecffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
]


def test_rfc9496_invalid_encodings_rejected(backend):
    import bpperm_b200
    for h in RFC9496_BAD_ENCODINGS:
        enc = bytes.fromhex(h)
        assert R.decompress(enc) is None, h  # the oracle agrees
        with pytest.raises(bpperm_b200.BppError) as ei:
            backend.upload_points(enc)
        assert ei.value.status == -5


def test_from_uniform_bytes_matches_oracle(backend):
    rng = ChaChaRng(b"\x11" * 32)
    blobs = [rng.fill_bytes(64) for _ in range(300)]
    blobs += [bytes(64), b"\xff" * 64, b"\x01" + bytes(63)]
    pts = backend.points_from_uniform(b"".join(blobs))
    enc = backend.compress_points(pts)
    for i, blob in enumerate(blobs):
        assert enc[32 * i:32 * i + 32] == R.compress(R.from_uniform_bytes(blob)), i
    import hashlib
    h = hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest()
    one = backend.compress_points(backend.points_from_uniform(h))
    assert one.hex() == "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"


def test_upload_formats_agree(backend):
    import struct
    pts = _points(50, 21)
    comp = b"".join(R.compress(p) for p in pts)
    aff = b"".join(_le(x) + _le(y) for x, y in (R.pt_affine(p) for p in pts))

    def radix51(v):
        return struct.pack("<5Q", *[(v >> (51 * i)) & ((1 << 51) - 1) for i in range(5)])

    # projective representatives with a random Z, as dalek keeps them in memory
    rnd = random.Random(5)
    xyzt = b""
    for p in pts:
        z = rnd.randrange(1, P)
        X, Y, Z, T = p[0] * z % P, p[1] * z % P, p[2] * z % P, p[3] * z % P
        xyzt += radix51(X) + radix51(Y) + radix51(Z) + radix51(T)
    for fmt, buf in ((0, comp), (1, aff), (2, xyzt)):
        h = backend.upload_points(buf, fmt)
        assert backend.compress_points(h) == comp, fmt
