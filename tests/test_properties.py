"""Property tests (hypothesis): size-independent identities the domain offers.

CPU: the "fold the scalars, not the generators" formulation the CUDA inner-product rounds use (ipa_kernels.cuh: s-table
by doubling, L_j / R_j as MSMs over the ORIGINAL generators) against the textbook generator folding of
bulletproofs 4.0.0 InnerProductProof::create as restated in oracle/ipa.py - for random vectors and challenges both must
give the same L_j, R_j and final a, b.  GPU: linearity of the MSM, batched == single, and sub-range consistency.
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import ristretto255 as R
from oracle.chacha import ChaChaRng

L = R.L


def _scalar_fold_rounds(G, H, hf, a, b, us):
    """L_j, R_j of every round computed WITHOUT folding generators: after r rounds
    s_t = prod_{k<r} (bit (r-1-k) of t ? u_k : u_k^-1); G^(r)[i] = sum_t s_t G[i + t n_r], H^(r)[i] = sum_t s_t^-1 hf.. H[..]."""
    n = len(G)
    lg = n.bit_length() - 1
    out = []
    s = [1]
    for r in range(lg):
        nj, h = n >> r, n >> (r + 1)
        mask = (1 << r) - 1
        vG, vH = [0] * n, [0] * n
        for g in range(n):
            i, t = g & (nj - 1), g >> (lg - r)
            ip = i ^ h
            vG[g] = a[ip] * s[t] % L
            vH[g] = b[ip] * s[mask - t] % L * hf[g] % L
        # L takes G's upper half-periods with a_lo and H's lower half-periods with b_hi; R the complement
        Ls, Lp, Rs, Rp = [], [], [], []
        for g in range(n):
            upper = (g & (nj - 1)) >= h
            (Ls if upper else Rs).append(vG[g]); (Lp if upper else Rp).append(G[g])
            (Rs if upper else Ls).append(vH[g]); (Rp if upper else Lp).append(H[g])
        out.append((R.msm_naive(Ls, Lp), R.msm_naive(Rs, Rp)))
        u, ui = us[r], R.sc_inv(us[r])
        a = [(a[i] * u + ui * a[h + i]) % L for i in range(h)] + a[h:]
        b = [(b[i] * ui + u * b[h + i]) % L for i in range(h)] + b[h:]
        s = [s[t >> 1] * (u if t & 1 else ui) % L for t in range(2 << r)]
    return out, a[0], b[0]


def _generator_fold_rounds(G, H, hf, a, b, us):
    """The textbook form (oracle/ipa.py InnerProductProof::create without transcript and c_L Q terms)."""
    n = len(G)
    G = list(G)
    H = [R.pt_mul(f, p) for f, p in zip(hf, H)]
    out = []
    r = 0
    while n != 1:
        n //= 2
        out.append((R.msm_naive(a[:n] + b[n:2 * n], G[n:2 * n] + H[:n]), R.msm_naive(a[n:2 * n] + b[:n], G[:n] + H[n:2 * n])))
        u, ui = us[r], R.sc_inv(us[r])
        a = [(a[i] * u + ui * a[n + i]) % L for i in range(n)]
        b = [(b[i] * ui + u * b[n + i]) % L for i in range(n)]
        G = [R.pt_add(R.pt_mul(ui, G[i]), R.pt_mul(u, G[n + i])) for i in range(n)]
        H = [R.pt_add(R.pt_mul(u, H[i]), R.pt_mul(ui, H[n + i])) for i in range(n)]
        r += 1
    return out, a[0], b[0]


@settings(max_examples=6, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(min_value=0, max_value=2**32 - 1), lg=st.integers(min_value=1, max_value=3))
def test_scalar_fold_equals_generator_fold(seed, lg):
    n = 1 << lg
    rng = ChaChaRng(seed.to_bytes(32, "little"))
    G = [rng.point() for _ in range(n)]
    H = [rng.point() for _ in range(n)]
    hf = [rng.scalar() for _ in range(n)]
    a = [rng.scalar() for _ in range(n)]
    b = [rng.scalar() for _ in range(n)]
    us = [rng.scalar() or 1 for _ in range(lg)]
    r1, a1, b1 = _scalar_fold_rounds(G, H, hf, list(a), list(b), us)
    r2, a2, b2 = _generator_fold_rounds(G, H, hf, list(a), list(b), us)
    assert (a1, b1) == (a2, b2)
    for (l1, rr1), (l2, rr2) in zip(r1, r2):
        assert R.compress(l1) == R.compress(l2) and R.compress(rr1) == R.compress(rr2)


@settings(max_examples=20, deadline=None)
@given(m=st.integers(min_value=1, max_value=130), salt=st.integers(min_value=0, max_value=255))
def test_commitment_digests_bind_every_commitment_and_their_order(m, salt):
    """the two-level V binding: changing, swapping, dropping or appending a commitment changes the digest list"""
    from oracle.acproof import commitment_digests
    rs = np.random.RandomState(m * 256 + salt)
    V = [rs.bytes(32) for _ in range(m)]
    d = commitment_digests(V)
    assert len(d) == (m + 63) // 64
    i = rs.randint(m)
    W = list(V)
    W[i] = bytes([W[i][0] ^ 1]) + W[i][1:]
    assert commitment_digests(W) != d
    if m > 1:
        j = (i + 1 + rs.randint(m - 1)) % m
        W = list(V)
        W[i], W[j] = W[j], W[i]
        assert (commitment_digests(W) != d) or V[i] == V[j]
    assert commitment_digests(V[:-1]) != d and commitment_digests(V + [V[0]]) != d


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(n=st.integers(min_value=1, max_value=700), seed=st.integers(min_value=0, max_value=2**31 - 1))
def test_msm_linearity_and_concatenation(backend, n, seed):
    """msm(a + b, P) == msm(a | b, P | P) and msm(a, P) over a split == msm of the halves' concatenation"""
    rs = np.random.RandomState(seed)
    blobs = rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes()
    a = [int.from_bytes(rs.bytes(32), "little") % L for _ in range(n)]
    b = [int.from_bytes(rs.bytes(32), "little") % L for _ in range(n)]
    sb = lambda v: b"".join(x.to_bytes(32, "little") for x in v)
    P = backend.points_from_uniform(blobs)
    PP = backend.points_from_uniform(blobs + blobs)
    lhs = backend.vartime_multiscalar_mul(sb([(x + y) % L for x, y in zip(a, b)]), P)
    rhs = backend.vartime_multiscalar_mul(sb(a) + sb(b), PP)
    assert lhs == rhs
    # c * msm(a, P) == msm(c a, P): the left side as a one-point MSM over the decompressed result
    c = int.from_bytes(rs.bytes(32), "little") % L
    base = backend.vartime_multiscalar_mul(sb(a), P)
    if base != bytes(32):
        assert backend.vartime_multiscalar_mul(sb([c]), [base]) == backend.vartime_multiscalar_mul(sb([c * x % L for x in a]), P)
    # the table path (precomputed points) agrees with the bucket path, single and batched
    backend.precompute(P, 7)
    assert backend.vartime_multiscalar_mul(sb(a), P) == base
    assert backend.msm_batch(sb(a) + sb(b), P, n, 2) == base + backend.vartime_multiscalar_mul(sb(b), PP, 0, n)
    P.free()
    PP.free()
