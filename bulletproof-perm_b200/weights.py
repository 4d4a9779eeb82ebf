"""Host-side circuit / witness / deck construction: the mirror of the reference's weights.rs.

`create_constants`, `create_a`, `create_weights`, `transpose` restate weights.rs:26-36,63-129,130-204 as
coded (they only make sense for 2-3 cards: SURVEY A.3 defects 6-9); `shuffle_circuit` / `shuffle_witness`
are the corrected k-card generator with the same shape (n = 2k multipliers, Q = 2n constraints,
m = 2k + 1 committed values) that the 52-card workload uses.  Pure host code (input generation).
"""
from __future__ import annotations

L = 2**252 + 27742317777372353535851937790883648493


def give_n(n: int) -> int:  # util.rs:96-103
    return n % L


def create_constants(Q: int):  # weights.rs:26-36
    return [0] * (Q - 2) + [L - 1, 1]


def transpose(v):  # weights.rs:115-129
    return [list(c) for c in zip(*v)]


def create_weights(card_count: int):  # weights.rs:130-204 (Q x n layout, as coded)
    n, Q = card_count * 2, card_count * 4
    w_l = [[0] * n for _ in range(Q)]
    w_r = [[0] * n for _ in range(Q)]
    w_o = [[0] * n for _ in range(Q)]
    w_v = [[0] * (n + 1) for _ in range(Q)]
    for i in range(Q):
        if i < n:
            w_l[i][i] = 1
            if i != card_count // 2 + 1 and i != 0:
                w_o[i][i - 1] = 1
            else:
                w_v[i][n] = L - 1
                w_v[i][i if i == 0 else i + 1] = 1
        else:
            w_r[i][i - n] = 1
            if i < Q - 2:
                w_v[i][n] = L - 1
                w_v[i][i - n + 1 if i < n + 3 else i - n + 2] = 1
    w_o[n - 1][card_count - 1] = 1
    return w_l, w_r, w_o, w_v


def create_a(variables):  # weights.rs:63-113 (as coded)
    n = len(variables) - 1
    a_L, a_R, a_O = [0] * n, [0] * n, [0] * n
    first, second, x = variables[:n // 2], variables[n // 2:n], variables[-1]
    offset = (n - 1) // 2
    for i in range(len(first) - 1):
        a_R[i] = (first[i + 1] - x) % L
        a_R[i + offset] = (second[i + 1] - x) % L
        if i == 0:
            a_L[i], a_L[i + offset] = (first[i] - x) % L, (second[i] - x) % L
        else:
            a_L[i], a_L[i + offset] = a_O[i - 1], a_O[i + offset - 1]
        a_O[i] = a_L[i] * a_R[i] % L
        a_O[i + offset] = a_L[i + offset] * a_R[i + offset] % L
    a_L[n - 2], a_R[n - 2] = a_O[n - 3], L - 1
    a_O[n - 2] = a_L[n - 2] * a_R[n - 2] % L
    a_L[n - 1], a_R[n - 1] = (a_O[offset] + a_O[n - 2]) % L, 1
    a_O[n - 1] = a_L[n - 1] * a_L[n - 1] % L
    return a_L, a_R, a_O


def shuffle_circuit(k: int):
    """prod_i (v_i - X) == prod_i (v_{k+i} - X) with X = v[2k]: two product chains of k-1 multipliers,
    one equality, two padding multipliers.  Returns (n, Q, m, W_L, W_R, W_O, W_V, c) with the W's as
    (wire, constraint, coeff) triples; constraint q: W_L a_L + W_R a_R + W_O a_O = W_V v + c."""
    if k < 2:
        raise ValueError("need at least two cards")
    n, Q, m = 2 * k, 4 * k, 2 * k + 1
    WL, WR, WO, WV = [], [], [], []
    q = 0
    for gb, vb in ((0, 0), (k - 1, k)):
        WL.append((gb, q, 1)); WV.append((vb, q, 1)); WV.append((2 * k, q, L - 1)); q += 1
        for i in range(k - 1):
            WR.append((gb + i, q, 1)); WV.append((vb + i + 1, q, 1)); WV.append((2 * k, q, L - 1)); q += 1
        for i in range(1, k - 1):
            WL.append((gb + i, q, 1)); WO.append((gb + i - 1, q, L - 1)); q += 1
    WO.append((k - 2, q, 1)); WO.append((2 * k - 3, q, L - 1)); q += 1
    WL.append((2 * k - 2, q, 1)); q += 1
    WL.append((2 * k - 1, q, 1)); q += 1
    return n, Q, m, WL, WR, WO, WV, [0] * Q


def shuffle_witness(k: int, perm, x: int):
    """deck 1..k (create_variables, weights.rs:38-56), `perm` a permutation of range(k), challenge value x.
    Returns (v, a_L, a_R, a_O)."""
    deck = [give_n(i) for i in range(1, k + 1)]
    v = deck + [deck[j] for j in perm] + [x % L]
    n = 2 * k
    a_L, a_R, a_O = [0] * n, [0] * n, [0] * n
    for gb, vb in ((0, 0), (k - 1, k)):
        for i in range(k - 1):
            a_L[gb + i] = (v[vb] - x) % L if i == 0 else a_O[gb + i - 1]
            a_R[gb + i] = (v[vb + i + 1] - x) % L
            a_O[gb + i] = a_L[gb + i] * a_R[gb + i] % L
    return v, a_L, a_R, a_O
