"""GPU parity: the reference's util.rs / poly.rs operators through the C ABI vs the oracle's
restatement (oracle/acproof.py).  Bit-exact (canonical 32-byte scalars)."""
import random

import pytest

from oracle import acproof as A, ristretto255 as R
from oracle.chacha import ChaChaRng

pytestmark = pytest.mark.gpu
L = R.L


@pytest.fixture(scope="module")
def ops(backend):
    import bpperm_b200
    return bpperm_b200.Ops(backend)


def _vec(rng, n):
    return [rng.scalar() for _ in range(n)]


EDGE = [0, 1, 2, L - 1, L - 2, (L - 1) // 2, 2**252, 2**128, 2**252 + 1]


@pytest.mark.parametrize("n", [1, 2, 31, 104, 208, 257, 1000])
def test_inner_product_and_hadamard(ops, n):
    rng = ChaChaRng(bytes([n % 256]) * 32)
    a, b = _vec(rng, n), _vec(rng, n)
    for i, e in enumerate(EDGE[:n]):
        a[i] = e
        b[-1 - i] = e
    assert ops.inner_product(a, b) == A.inner_product(a, b)
    assert ops.hadamard_V(a, b) == A.hadamard_V(a, b)


def test_dimension_mismatch_is_an_error(ops):
    rng = ChaChaRng(b"\x01" * 32)
    a, b = _vec(rng, 4), _vec(rng, 5)
    with pytest.raises(ValueError):
        ops.inner_product(a, b)          # util.rs:86-88 panics
    with pytest.raises(ValueError):
        ops.hadamard_V(a, b)             # util.rs:9-11
    with pytest.raises(ValueError):
        ops.vm_mult(a, [b, b])           # util.rs:26-28
    with pytest.raises(ValueError):
        ops.mv_mult([a, a, a], b)        # util.rs:44-46
    assert ops.inner_product([], []) == 0


def test_vm_mult_mv_mult_dense_52_card_shapes(ops):
    # W_L is n x Q = 104 x 208, W_V is m x Q = 105 x 208; 0/+-1 entries like the circuit
    n, Q, m, WL, WR, WO, WV, c = A.shuffle_circuit(52)
    rng = ChaChaRng(b"\x02" * 32)
    z_q = A.exp_iter(rng.scalar(), Q)
    for triples, rows in ((WL, n), (WO, n), (WV, m)):
        M = A.dense(triples, rows, Q)
        assert ops.vm_mult(z_q, M) == A.vm_mult(z_q, M)
        v = _vec(rng, rows)
        assert ops.mv_mult(M, v) == A.mv_mult(M, v)
    # fully random small matrix
    M = [_vec(rng, 7) for _ in range(5)]
    assert ops.vm_mult(_vec(ChaChaRng(b"\x03" * 32), 7), M) == A.vm_mult(_vec(ChaChaRng(b"\x03" * 32), 7), M)
    v = _vec(rng, 5)
    assert ops.mv_mult(M, v) == A.mv_mult(M, v)
    assert ops.lm_mult(tuple(z_q), A.dense(WL, n, Q)) == A.lm_mult(tuple(z_q), A.dense(WL, n, Q))


def test_exp_iter_reproduces_the_reference_fibonacci_exponents(ops):
    # with y from the E.3 transcript, the 8th output of exp_iter is y^21
    y = int.from_bytes(bytes.fromhex("134ddbf9759905621d6baea3534e91e67a9ed02d27571d58f1dac13b401ad000"), "little")
    got = ops.exp_iter(y, 208)
    assert got == A.exp_iter(y, 208)
    assert got[:8] == [pow(y, e, L) for e in (1, 1, 2, 3, 5, 8, 13, 21)]
    assert R.sc_bytes(got[7]).hex() == "8af006d688850f521ca0d7b6460ef02f9cc4dbb59b87e7f9f928568f2650ac0c"
    assert ops.exp_iter(0, 3) == [0, 0, 0] and ops.exp_iter(1, 3) == [1, 1, 1]


def test_scalar_powers_exp_invert(ops):
    rng = ChaChaRng(b"\x04" * 32)
    x = rng.scalar()
    assert ops.scalar_powers(x, 0, 300) == [pow(x, i, L) for i in range(300)]
    for p in (0, 1, 2, 3, 6):
        assert ops.scalar_exp(x, p) == A.scalar_exp(x, p)
    xs = _vec(rng, 104) + EDGE
    assert ops.invert_all(xs) == [R.sc_inv(v) for v in xs]   # 0 -> 0 like dalek's x^(l-2)


def test_scalar_invert_binary_gcd_worst_cases(ops):
    """Scalar::invert (circuit_lib.rs:273-275) is a fixed-length binary extended GCD on the device: powers of two,
    their neighbours and negatives take the most steps (505 of the 508 performed); 0 -> 0 like dalek's x^(l-2)."""
    import random
    rnd = random.Random(77)
    xs = [0, 1, 2, 3, L - 1, L - 2, (L - 1) // 2, (L + 1) // 2]
    for k in range(1, 253):
        xs += [pow(2, k, L), (pow(2, k, L) - 1) % L, (L - pow(2, k, L)) % L]
    xs += [rnd.randrange(L) for _ in range(2000)]
    got = ops.invert_all(xs)
    for x, g in zip(xs, got):
        assert g == (pow(x, L - 2, L)), hex(x)


def test_wide_reduce_and_reduce(ops):
    rng = ChaChaRng(b"\x05" * 32)
    blobs = [rng.fill_bytes(64) for _ in range(200)] + [bytes(64), b"\xff" * 64, b"\x00" * 32 + b"\xff" * 32]
    assert ops.from_bytes_mod_order_wide(b"".join(blobs)) == [R.sc_from_wide(b) for b in blobs]
    # SURVEY E.1
    z = bytes.fromhex("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                      "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    assert R.sc_bytes(ops.from_bytes_mod_order_wide(z)[0]).hex() == "4a53c3fbbc59970ee5f85af813875dffc13a904a2e53ae7e65fa0dea6e62c901"
    raw = [rng.fill_bytes(32) for _ in range(50)] + [b"\xff" * 32, (L).to_bytes(32, "little"), (L - 1).to_bytes(32, "little")]
    assert ops.reduce_scalars(b"".join(raw)) == [int.from_bytes(r, "little") % L for r in raw]


@pytest.mark.parametrize("n", [1, 6, 104, 513])
def test_vecpoly3_and_poly6(ops, n):
    rng = ChaChaRng(bytes([n % 256, 9]) * 16)
    lx, rx = A.VecPoly3(n), A.VecPoly3(n)
    lx.c = [_vec(rng, n) for _ in range(4)]
    rx.c = [_vec(rng, n) for _ in range(4)]
    lx.c[0] = [0] * n      # as in the protocol: l0 = 0, r2 = 0 (circuit_lib.rs:313-339)
    rx.c[2] = [0] * n
    want = A.VecPoly3.special_inner_product(lx, rx)
    t = ops.special_inner_product(lx.c, rx.c)
    assert t == want.t
    x = rng.scalar()
    assert ops.vecpoly3_eval(lx.c, x) == lx.eval(x)
    assert ops.poly6_eval(t, x) == want.eval(x)
    for i in (1, 3, 4, 5, 6):   # the reference evaluates t(X) at small integers (circuit_lib.rs:362-406)
        assert ops.poly6_eval(t, i) == want.eval(i)
