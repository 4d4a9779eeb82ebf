"""Restatement of the inner-product argument and of the `fixed` protocol mode that uses it.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference has NO inner-product argument: `circuit_lib.rs:464-468` sends l and r in the clear and
`lib.rs:30` only imports the `bulletproofs` crate (4.0.0, Cargo.lock:47-50; un-vendored) for its
generators.  SURVEY section 8 row a16 / north_star item (2) ask for `InnerProductProof::create` folding
anyway, so this module restates the published algorithm of bulletproofs 4.0.0
`src/inner_product_proof.rs` (create / verification_scalars / verify) and the glue dalek's R1CS prover
puts around it (`src/r1cs/prover.rs`, `verifier.rs`: the "t_x", "t_x_blinding", "e_blinding" appends,
the challenge "w", Q = w*B, H_factors = y^-n, padding to a power of two).  PARITY UNPINNED for this
part: there is no reference byte string and no published vector; parity = this file == oracle/c ==
the CUDA path, byte for byte.

`fixed` mode = `reference-fixed` (oracle/acproof.py) with
  * standard powers y^0..y^(n'-1) and z^1..z^Q instead of exp_iter's Fibonacci exponents (defect 1),
  * l, r padded to n' = next power of two (l: 0, r: -y^i, as dalek pads) and replaced in the proof by
    the inner-product proof (L_j, R_j, a, b).
Proof bytes: A_I, A_O, S, T_1, T_3, T_4, T_5, T_6 | t_hat, tau_x, mu | L_0, R_0, .., L_{k-1}, R_{k-1} | a, b.
"""
from __future__ import annotations

from . import ristretto255 as R
from .merlin import Transcript
from .acproof import (VecPoly3, append_commitments, commit_variables, hadamard_V, inner_product, mv_mult, scalar_exp,  # noqa: F401
                      vm_mult)

L = R.L


def next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def std_powers(x: int, count: int, first: int = 1):
    """bulletproofs util::exp_iter (the intended behaviour): first, first*x, first*x^2, ..."""
    out, cur = [], first % L
    for _ in range(count):
        out.append(cur)
        cur = cur * x % L
    return out


def innerproduct_domain_sep(trans: Transcript, n: int):  # bulletproofs transcript.rs
    trans.append_message(b"dom-sep", b"ipp v1")
    trans.append_u64(b"n", n)


class InnerProductProof:
    def __init__(self, L_vec, R_vec, a, b):
        self.L_vec, self.R_vec, self.a, self.b = L_vec, R_vec, a, b

    # inner_product_proof.rs: InnerProductProof::create
    @classmethod
    def create(cls, trans: Transcript, Q, G_factors, H_factors, G_vec, H_vec, a_vec, b_vec, msm=None):
        msm = msm or R.vartime_multiscalar_mul
        n = len(G_vec)
        assert n & (n - 1) == 0 and n >= 1
        assert len(H_vec) == n and len(a_vec) == n and len(b_vec) == n
        assert len(G_factors) == n and len(H_factors) == n
        G, H, a, b = list(G_vec), list(H_vec), list(a_vec), list(b_vec)
        innerproduct_domain_sep(trans, n)
        L_vec, R_vec = [], []
        first = True
        while n != 1:
            n //= 2
            a_L, a_R = a[:n], a[n:]
            b_L, b_R = b[:n], b[n:]
            G_L, G_R = G[:n], G[n:]
            H_L, H_R = H[:n], H[n:]
            c_L = inner_product(a_L, b_R)
            c_R = inner_product(a_R, b_L)
            if first:   # the first round carries the generator factors
                gfl, gfr = G_factors[:n], G_factors[n:2 * n]
                hfl, hfr = H_factors[:n], H_factors[n:2 * n]
            else:
                gfl = gfr = hfl = hfr = [1] * n
            Lp = msm([x * f % L for x, f in zip(a_L, gfr)] + [x * f % L for x, f in zip(b_R, hfl)] + [c_L],
                     G_R + H_L + [Q])
            Rp = msm([x * f % L for x, f in zip(a_R, gfl)] + [x * f % L for x, f in zip(b_L, hfr)] + [c_R],
                     G_L + H_R + [Q])
            Lc, Rc = R.compress(Lp), R.compress(Rp)
            L_vec.append(Lc)
            R_vec.append(Rc)
            trans.append_point(b"L", Lc)
            trans.append_point(b"R", Rc)
            u = trans.challenge_scalar(b"u")
            u_inv = R.sc_inv(u)
            a = [(x * u + u_inv * y) % L for x, y in zip(a_L, a_R)]
            b = [(x * u_inv + u * y) % L for x, y in zip(b_L, b_R)]
            G = [msm([u_inv * fl % L, u * fr % L], [p, q]) for p, q, fl, fr in zip(G_L, G_R, gfl, gfr)]
            H = [msm([u * fl % L, u_inv * fr % L], [p, q]) for p, q, fl, fr in zip(H_L, H_R, hfl, hfr)]
            first = False
        return cls(L_vec, R_vec, a[0], b[0])

    # inner_product_proof.rs: verification_scalars -> (u_sq, u_inv_sq, s)
    def verification_scalars(self, n: int, trans: Transcript):
        lg_n = len(self.L_vec)
        if lg_n >= 32 or n != (1 << lg_n):
            return None
        innerproduct_domain_sep(trans, n)
        chal = []
        for Lc, Rc in zip(self.L_vec, self.R_vec):
            if not trans.validate_and_append_point(b"L", Lc):
                return None
            if not trans.validate_and_append_point(b"R", Rc):
                return None
            chal.append(trans.challenge_scalar(b"u"))
        chal_inv = [R.sc_inv(u) for u in chal]
        allinv = 1
        for ui in chal_inv:
            allinv = allinv * ui % L
        u_sq = [u * u % L for u in chal]
        u_inv_sq = [u * u % L for u in chal_inv]
        s = [allinv]
        for i in range(1, n):
            lg_i = i.bit_length() - 1
            k = 1 << lg_i
            s.append(s[i - k] * u_sq[(lg_n - 1) - lg_i] % L)
        return u_sq, u_inv_sq, s

    def to_bytes(self) -> bytes:
        out = b"".join(l + r for l, r in zip(self.L_vec, self.R_vec))
        return out + R.sc_bytes(self.a) + R.sc_bytes(self.b)


def proof_len(n: int) -> int:
    lg = next_pow2(n).bit_length() - 1
    return 32 * (13 + 2 * lg)


def prove(core, prover, V, rng, label=b"test", msm=None):
    """`fixed`-mode prover.  core["G_vec"], core["H_vec"] must hold at least next_pow2(n) generators; V = the value
    commitments (bound to the transcript right after the domain separator, oracle.acproof.append_commitments).
    Returns (proof_bytes, state) - state carries the intermediate values tests look at."""
    msm = msm or R.vartime_multiscalar_mul
    n, Q, m = core["n"], core["Q"], core["m"]
    npad = next_pow2(n)
    g, h = core["g_base"], core["h_base"]
    G, H = core["G_vec"][:npad], core["H_vec"][:npad]
    assert len(G) == npad and len(H) == npad
    a_L, a_R, a_O, gamma = prover["a_L"], prover["a_R"], prover["a_O"], prover["gamma"]
    trans = Transcript(label)
    trans.arithmetic_domain_sep(n)
    append_commitments(trans, V, m)
    alpha, beta, ro = rng.scalar(), rng.scalar(), rng.scalar()
    A_I = msm([alpha] + a_L + a_R, [h] + G[:n] + H[:n])
    A_O = msm([beta] + a_O, [h] + G[:n])
    s_l = [rng.scalar() for _ in range(n)]
    s_r = [rng.scalar() for _ in range(n)]
    S = msm([ro] + s_l + s_r, [h] + G[:n] + H[:n])
    pts = [R.compress(A_I), R.compress(A_O), R.compress(S)]
    for lab, p in zip((b"A_I", b"A_O", b"S"), pts):
        trans.append_point(lab, p)
    y = trans.challenge_scalar(b"y")
    z = trans.challenge_scalar(b"z")
    y_n = std_powers(y, npad)
    y_inv = R.sc_inv(y)
    y_n_inv = std_powers(y_inv, npad)
    z_q = std_powers(z, Q, z)
    zWL, zWR, zWO = vm_mult(z_q, core["W_L"]), vm_mult(z_q, core["W_R"]), vm_mult(z_q, core["W_O"])
    l_in = hadamard_V(y_n_inv[:n], zWR)
    l_x, r_x = VecPoly3(n), VecPoly3(n)
    l_x.c[1] = [(a + b) % L for a, b in zip(a_L, l_in)]
    l_x.c[2] = list(a_O)
    l_x.c[3] = list(s_l)
    r_x.c[0] = [(a - b) % L for a, b in zip(zWO, y_n[:n])]
    r_x.c[1] = [(a + b) % L for a, b in zip(hadamard_V(y_n[:n], a_R), zWL)]
    r_x.c[3] = hadamard_V(y_n[:n], s_r)
    t_poly = VecPoly3.special_inner_product(l_x, r_x)
    taus = []
    for deg, lab in ((1, b"T1"), (3, b"T3"), (4, b"T4"), (5, b"T5"), (6, b"T6")):
        tau = rng.scalar()
        taus.append(tau)
        T = R.compress(msm([t_poly.t[deg - 1], tau], [g, h]))
        pts.append(T)
        trans.append_point(lab, T)
    x = trans.challenge_scalar(b"x")
    l = l_x.eval(x) + [0] * (npad - n)
    r = r_x.eval(x) + [(L - y_n[i]) % L for i in range(n, npad)]
    t_hat = inner_product(l, r)
    xx = x * x % L
    tau_x = xx * inner_product(z_q, mv_mult(core["W_V"], gamma)) % L
    for tau, deg in zip(taus, (1, 3, 4, 5, 6)):
        tau_x = (tau_x + tau * scalar_exp(x, deg)) % L
    mu = (alpha * x + beta * xx + ro * xx * x) % L
    trans.append_scalar(b"t_x", t_hat)
    trans.append_scalar(b"t_x_blinding", tau_x)
    trans.append_scalar(b"e_blinding", mu)
    w = trans.challenge_scalar(b"w")
    Qp = R.pt_mul(w, g)
    ipp = InnerProductProof.create(trans, Qp, [1] * npad, y_n_inv, G, H, l, r, msm)
    proof = b"".join(pts) + R.sc_bytes(t_hat) + R.sc_bytes(tau_x) + R.sc_bytes(mu) + ipp.to_bytes()
    assert len(proof) == proof_len(n)
    return proof, {"y": y, "z": z, "x": x, "w": w, "l": l, "r": r, "t_hat": t_hat, "ipp": ipp}


def verify(core, V, proof: bytes, label=b"test", msm=None) -> bool:
    """`fixed`-mode verifier, un-fused: check 2 (circuit_lib.rs:521-544) and check 3 (:551-575) with the
    right-hand side <l,G> + <r,h'> replaced by the inner-product verification.  Scalars in the proof are
    taken mod l; an invalid or identity L/R encoding, or an invalid point, rejects."""
    msm = msm or R.vartime_multiscalar_mul
    n, Q, m = core["n"], core["Q"], core["m"]
    npad = next_pow2(n)
    lg = npad.bit_length() - 1
    if len(proof) != proof_len(n):
        return False
    g, h = core["g_base"], core["h_base"]
    G, H = core["G_vec"][:npad], core["H_vec"][:npad]
    w32 = [proof[32 * i:32 * i + 32] for i in range(len(proof) // 32)]
    pts = [R.decompress(b) for b in w32[:8]]
    if any(p is None for p in pts):
        return False
    A_I, A_O, S = pts[:3]
    T_pts = pts[3:8]
    t_hat, tau_x, mu = (R.sc_from_bytes_mod_order(b) for b in w32[8:11])
    ipp = InnerProductProof([w32[11 + 2 * j] for j in range(lg)], [w32[12 + 2 * j] for j in range(lg)],
                            R.sc_from_bytes_mod_order(w32[11 + 2 * lg]), R.sc_from_bytes_mod_order(w32[12 + 2 * lg]))
    trans = Transcript(label)
    trans.arithmetic_domain_sep(n)
    if len(V) != m:
        return False
    append_commitments(trans, V, m)
    for lab, b in zip((b"A_I", b"A_O", b"S"), w32[:3]):
        trans.append_point(lab, b)
    y = trans.challenge_scalar(b"y")
    z = trans.challenge_scalar(b"z")
    for lab, b in zip((b"T1", b"T3", b"T4", b"T5", b"T6"), w32[3:8]):
        trans.append_point(lab, b)
    x = trans.challenge_scalar(b"x")
    trans.append_scalar(b"t_x", t_hat)
    trans.append_scalar(b"t_x_blinding", tau_x)
    trans.append_scalar(b"e_blinding", mu)
    w = trans.challenge_scalar(b"w")
    vs = ipp.verification_scalars(npad, trans)
    if vs is None:
        return False
    u_sq, u_inv_sq, s = vs
    L_pts = [R.decompress(b) for b in ipp.L_vec]
    R_pts = [R.decompress(b) for b in ipp.R_vec]
    if any(p is None for p in L_pts + R_pts):
        return False
    y_n = std_powers(y, npad)
    y_n_inv = std_powers(R.sc_inv(y), npad)
    z_q = std_powers(z, Q, z)
    zWL, zWR, zWO = vm_mult(z_q, core["W_L"]), vm_mult(z_q, core["W_R"]), vm_mult(z_q, core["W_O"])
    zWV = vm_mult(z_q, core["W_V"])
    l_in = hadamard_V(y_n_inv[:n], zWR)
    sigma = inner_product(l_in, zWL)
    xx = x * x % L
    # check 2: t_hat*g + tau_x*h == x^2(<z_q,c> + sigma)*g + sum_j x^2 (z W_V)_j V_j + sum_i x^i T_i
    g_exp = xx * (inner_product(z_q, core["c_vec"]) + sigma) % L
    cand = msm([g_exp] + [xx * v % L for v in zWV] + [scalar_exp(x, d) for d in (1, 3, 4, 5, 6)], [g] + list(V) + T_pts)
    if not R.pt_eq(msm([t_hat, tau_x], [g, h]), cand):
        return False
    # check 3: P = x A_I + x^2 A_O + x^3 S - mu h + <x l_in, G> + <y^-n o (x zWL + zWO - y^n), H>  (= <l,G> + <r,h'>)
    hs = []
    for i in range(npad):
        inner = ((x * zWL[i] + zWO[i]) if i < n else 0) - y_n[i]
        hs.append(y_n_inv[i] * inner % L)
    P = msm([x, xx, xx * x % L, (L - mu) % L] + [x * v % L for v in l_in] + hs, [A_I, A_O, S, h] + G[:n] + H)
    # inner-product verification (InnerProductProof::verify) of P + t_hat*Q with Q = w*g
    a, b = ipp.a, ipp.b
    Qp = R.pt_mul(w, g)
    gs = [a * si % L for si in s]
    hsc = [b * s[npad - 1 - i] % L * y_n_inv[i] % L for i in range(npad)]
    rhs = msm(gs + hsc + [a * b % L] + [(L - v) % L for v in u_sq] + [(L - v) % L for v in u_inv_sq],
              G + H + [Qp] + L_pts + R_pts)
    return R.pt_eq(R.pt_add(P, R.pt_mul(t_hat, Qp)), rhs)


def make_instance(k, rng, dense_weights=True):
    """oracle.acproof.make_instance with next_pow2(n) generators (same RNG draw order otherwise)."""
    from . import acproof as A
    g, h = rng.point(), rng.point()
    n, Q, m, WL, WR, WO, WV, c = A.shuffle_circuit(k)
    npad = next_pow2(n)
    G = [rng.point() for _ in range(npad)]
    H = [rng.point() for _ in range(npad)]
    v, a_L, a_R, a_O = A.shuffle_witness(k, rng)
    gamma = [rng.scalar() for _ in range(m)]
    V = commit_variables(v, gamma, g, h)
    core = {"g_base": g, "h_base": h, "G_vec": G, "H_vec": H, "c_vec": c, "sparse": (WL, WR, WO, WV),
            "n": n, "Q": Q, "m": m}
    if dense_weights:
        core.update(W_L=A.dense(WL, n, Q), W_R=A.dense(WR, n, Q), W_O=A.dense(WO, n, Q), W_V=A.dense(WV, m, Q))
    prover = {"a_L": a_L, "a_R": a_R, "a_O": a_O, "gamma": gamma, "v": v}
    return core, prover, V
