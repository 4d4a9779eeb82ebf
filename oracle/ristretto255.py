"""Ristretto255 / Edwards25519 / scalar-mod-l oracle in Python big integers.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates the published
algorithms of curve25519-dalek-ng 4.1.1 (pinned at
/root/reference/bp-perm/Cargo.lock:109-112, NOT vendored) == RFC 9496, which
every group operation of the reference goes through:

* `RistrettoPoint::vartime_multiscalar_mul`  - 15 call sites,
  /root/reference/bp-perm/src/circuit_lib.rs:187,202,216,363,374,385,396,407,
  498,504,509,525,535,552,568
* `RistrettoPoint * Scalar`                  - circuit_lib.rs:491
* `.compress()` / `.decompress()` / `!=`     - circuit_lib.rs:231-233,368,532,541
* `RistrettoPoint::random`                   - lib.rs:165-166,179-180
* `Scalar::{random,invert,from_bytes_mod_order_wide}` - circuit_lib.rs:180-182,
  273-275; transcript_protocol.rs:62-67

Points are tuples (X, Y, Z, T) of ints mod p in extended twisted Edwards
coordinates (a = -1).  Scalars are ints in [0, l).
"""
from __future__ import annotations

import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
D2 = (2 * D) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
if SQRT_M1 & 1:  # dalek / RFC 9496 use the even ("non-negative") root
    SQRT_M1 = P - SQRT_M1
assert (SQRT_M1 * SQRT_M1) % P == P - 1


def _is_neg(x: int) -> int:
    return (x % P) & 1


def _abs(x: int) -> int:
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_i(u: int, v: int):
    """RFC 9496 4.2 SQRT_RATIO_M1 == dalek FieldElement::sqrt_ratio_i."""
    u %= P
    v %= P
    v3 = (v * v * v) % P
    v7 = (v3 * v3 * v) % P
    r = (u * v3 * pow(u * v7, (P - 5) // 8, P)) % P
    check = (v * r * r) % P
    correct = check == u
    flipped = check == (P - u) % P
    flipped_i = check == ((P - u) * SQRT_M1) % P
    if flipped or flipped_i:
        r = (r * SQRT_M1) % P
    return (correct or flipped), _abs(r)


def _const_sqrt(x):
    ok, r = sqrt_ratio_i(x % P, 1)
    assert ok
    return r


# RFC 9496 4.1 constants, derived and checked against their defining equations.
INVSQRT_A_MINUS_D = sqrt_ratio_i(1, (-1 - D) % P)[1]
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = ((D - 1) * (D - 1)) % P
assert (SQRT_AD_MINUS_ONE * SQRT_AD_MINUS_ONE) % P == (-D - 1) % P
assert (INVSQRT_A_MINUS_D * INVSQRT_A_MINUS_D * (-1 - D)) % P == 1
assert INVSQRT_A_MINUS_D == 54469307008909316920995813868745141605393597292927456921205312896311721017578

IDENTITY = (0, 1, 1, 0)

# RFC 9496 / dalek RISTRETTO_BASEPOINT (the ed25519 base point)
_BY = (4 * pow(5, P - 2, P)) % P
_BX = sqrt_ratio_i((_BY * _BY - 1) % P, (D * _BY * _BY + 1) % P)[1]
BASEPOINT = (_BX, _BY, 1, (_BX * _BY) % P)


# ----------------------------------------------------------------- Edwards --
def pt_add(p, q):
    """add-2008-hwcd-3 (dalek EdwardsPoint + ProjectiveNiels, curve_models)."""
    X1, Y1, Z1, T1 = p
    X2, Y2, Z2, T2 = q
    A = ((Y1 - X1) * (Y2 - X2)) % P
    B = ((Y1 + X1) * (Y2 + X2)) % P
    C = (T1 * D2 % P) * T2 % P
    Dd = (2 * Z1 * Z2) % P
    E, F, G, H = B - A, Dd - C, Dd + C, B + A
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def pt_neg(p):
    X, Y, Z, T = p
    return ((-X) % P, Y, Z, (-T) % P)


def pt_sub(p, q):
    return pt_add(p, pt_neg(q))


def pt_double(p):
    """dbl-2008-hwcd (dalek ProjectivePoint::double)."""
    X1, Y1, Z1, _ = p
    A = X1 * X1 % P
    B = Y1 * Y1 % P
    C = 2 * Z1 * Z1 % P
    H = A + B
    E = H - (X1 + Y1) * (X1 + Y1) % P
    G = A - B
    F = C + G
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def pt_mul(k: int, p):
    """Plain double-and-add; the group result is algorithm independent."""
    k %= L
    acc = IDENTITY
    for bit in bin(k)[2:] if k else "":
        acc = pt_double(acc)
        if bit == "1":
            acc = pt_add(acc, p)
    return acc


def pt_eq(p, q) -> bool:
    """Ristretto equality (RFC 9496 4.3.3): X1Y2==Y1X2 or X1X2==Y1Y2."""
    X1, Y1, _, _ = p
    X2, Y2, _, _ = q
    return (X1 * Y2 - Y1 * X2) % P == 0 or (X1 * X2 - Y1 * Y2) % P == 0


def pt_is_identity(p) -> bool:
    return pt_eq(p, IDENTITY)


def pt_affine(p):
    X, Y, Z, _ = p
    zi = pow(Z, P - 2, P)
    return (X * zi % P, Y * zi % P)


# --------------------------------------------------------------- Ristretto --
def compress(p) -> bytes:
    """RFC 9496 4.3.2 Encode == dalek RistrettoPoint::compress."""
    X, Y, Z, T = p
    u1 = (Z + Y) * (Z - Y) % P
    u2 = X * Y % P
    _, inv = sqrt_ratio_i(1, u1 * u2 * u2 % P)
    i1 = inv * u1 % P
    i2 = inv * u2 % P
    z_inv = i1 * i2 % P * T % P
    den_inv = i2
    if _is_neg(T * z_inv):
        X, Y = Y * SQRT_M1 % P, X * SQRT_M1 % P
        den_inv = i1 * INVSQRT_A_MINUS_D % P
    if _is_neg(X * z_inv):
        Y = (-Y) % P
    s = _abs(den_inv * (Z - Y))
    return s.to_bytes(32, "little")


def decompress(b: bytes):
    """RFC 9496 4.3.1 Decode == dalek CompressedRistretto::decompress.
    Returns None for invalid encodings."""
    if len(b) != 32:
        return None
    s = int.from_bytes(b, "little")
    if s >= P or (s & 1):
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2s = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2s) % P
    ok, inv = sqrt_ratio_i(1, v * u2s % P)
    dx = inv * u2 % P
    dy = inv * dx % P * v % P
    x = _abs(2 * s * dx)
    y = u1 * dy % P
    t = x * y % P
    if (not ok) or _is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def elligator(r0: int):
    """RFC 9496 4.3.4 MAP == dalek RistrettoPoint::elligator_ristretto_flavor."""
    r = SQRT_M1 * r0 % P * r0 % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) % P * ((r + D) % P) % P
    was_sq, s = sqrt_ratio_i(u, v)
    s_prime = (-_abs(s * r0)) % P
    if not was_sq:
        s = s_prime
        c = r
    else:
        c = P - 1
    N = (c * (r - 1) % P * D_MINUS_ONE_SQ - v) % P
    w0 = 2 * s * v % P
    w1 = N * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return (w0 * w3 % P, w2 * w1 % P, w1 * w3 % P, w0 * w2 % P)


def from_uniform_bytes(b: bytes):
    """dalek RistrettoPoint::from_uniform_bytes == RFC 9496 one-way map
    (== libsodium crypto_core_ristretto255_from_hash)."""
    assert len(b) == 64
    r1 = int.from_bytes(b[:32], "little") & ((1 << 255) - 1)
    r2 = int.from_bytes(b[32:], "little") & ((1 << 255) - 1)
    return pt_add(elligator(r1 % P), elligator(r2 % P))


def hash_to_point_sha3_512(msg: bytes):
    """dalek RistrettoPoint::hash_from_bytes::<Sha3_512> (PedersenGens::default
    B_blinding, bulletproofs 4.0.0 generators.rs)."""
    return from_uniform_bytes(hashlib.sha3_512(msg).digest())


# ----------------------------------------------------------------- scalars --
def sc_from_wide(b: bytes) -> int:
    """Scalar::from_bytes_mod_order_wide (transcript_protocol.rs:62-67)."""
    assert len(b) == 64
    return int.from_bytes(b, "little") % L


def sc_from_bytes_mod_order(b: bytes) -> int:
    return int.from_bytes(b, "little") % L


def sc_bytes(s: int) -> bytes:
    return (s % L).to_bytes(32, "little")


def sc_inv(s: int) -> int:
    return pow(s % L, L - 2, L)


# ---------------------------------------------------------------- MSM ------
def msm_naive(scalars, points):
    acc = IDENTITY
    for s, p in zip(scalars, points):
        acc = pt_add(acc, pt_mul(s, p))
    return acc


def to_radix_2w(s: int, w: int):
    """dalek Scalar::to_radix_2w(w) for 4 <= w <= 8 (scalar.rs), generalised to
    any w: signed digits in [-2^(w-1), 2^(w-1)]; the last carry is folded the
    way dalek does (extra digit for w == 8, into the top digit otherwise)."""
    radix = 1 << w
    mask = radix - 1
    digits_count = (256 + w - 1) // w
    digits = []
    carry = 0
    for i in range(digits_count):
        coef = carry + ((s >> (w * i)) & mask)
        carry = (coef + (radix >> 1)) >> w
        digits.append(coef - (carry << w))
    if w == 8:
        digits.append(carry)
    else:
        digits[-1] += carry << w
    return digits


def non_adjacent_form(s: int, w: int):
    """dalek Scalar::non_adjacent_form(w): 256 signed odd digits |d| < 2^(w-1)."""
    naf = [0] * 256
    width = 1 << w
    window_mask = width - 1
    pos = 0
    carry = 0
    while pos < 256:
        bit_buf = (s >> pos) & window_mask if pos < 256 else 0
        window = carry + bit_buf
        if window & 1 == 0:
            pos += 1
            continue
        if window < width // 2:
            carry = 0
            naf[pos] = window
        else:
            carry = 1
            naf[pos] = window - width
        pos += w
    return naf


def msm_straus(scalars, points):
    """dalek backend::serial::scalar_mul::straus::Straus vartime
    (optional_multiscalar_mul, NAF width 5)."""
    nafs = [non_adjacent_form(s % L, 5) for s in scalars]
    tables = []
    for p in points:
        p2 = pt_double(p)
        t = [p]
        for _ in range(7):
            t.append(pt_add(t[-1], p2))
        tables.append(t)  # [P, 3P, ..., 15P]
    r = IDENTITY
    for i in range(255, -1, -1):
        r = pt_double(r)
        for naf, tab in zip(nafs, tables):
            d = naf[i]
            if d > 0:
                r = pt_add(r, tab[d // 2])
            elif d < 0:
                r = pt_sub(r, tab[(-d) // 2])
    return r


def msm_pippenger(scalars, points, w=None):
    """dalek backend::serial::scalar_mul::pippenger::Pippenger vartime."""
    size = len(scalars)
    if w is None:
        w = 6 if size < 500 else (7 if size < 800 else 8)
    max_digit = 1 << w
    digits_count = (256 + w - 1) // w + (1 if w == 8 else 0)
    buckets_count = max_digit // 2
    sd = [to_radix_2w(s % L, w) for s in scalars]

    def column(di):
        buckets = [IDENTITY] * buckets_count
        for digs, pt in zip(sd, points):
            d = digs[di]
            if d > 0:
                buckets[d - 1] = pt_add(buckets[d - 1], pt)
            elif d < 0:
                buckets[-d - 1] = pt_sub(buckets[-d - 1], pt)
        inter = buckets[buckets_count - 1]
        total = buckets[buckets_count - 1]
        for i in range(buckets_count - 2, -1, -1):
            inter = pt_add(inter, buckets[i])
            total = pt_add(total, inter)
        return total

    hi = column(digits_count - 1)
    for di in range(digits_count - 2, -1, -1):
        for _ in range(w):
            hi = pt_double(hi)
        hi = pt_add(hi, column(di))
    return hi


def vartime_multiscalar_mul(scalars, points):
    """RistrettoPoint::vartime_multiscalar_mul with dalek's size dispatch
    (EdwardsPoint::optional_multiscalar_mul: < 190 Straus, else Pippenger)."""
    scalars = list(scalars)
    points = list(points)
    assert len(scalars) == len(points)  # dalek asserts equal exact size hints
    if len(scalars) < 190:
        return msm_straus(scalars, points)
    return msm_pippenger(scalars, points)
