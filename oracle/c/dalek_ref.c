/* dalek_ref.c - CPU restatement of the arithmetic under bulletproof-perm's hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as the fast checker, by
 * bench.py as the timed CPU baseline ("C restatement of curve25519-dalek-ng 4.1.1 serial u64
 * backend" - never "dalek") and by __graft_entry__.smoke().  The product never links it.
 *
 * The reference (/root/reference/bp-perm) does all group arithmetic through curve25519-dalek-ng
 * 4.1.1 (Cargo.lock:109-112; serial u64 backend, no SIMD feature), which is not vendored.  This
 * file restates that crate's published algorithms with the same data representation and the
 * same algorithmic choices, so that its timing is a like-for-like stand-in:
 *   - FieldElement51: 5 x 51-bit limbs, u128 products            (backend/serial/u64/field.rs)
 *   - EdwardsPoint / ProjectiveNiels / AffineNiels / Completed    (backend/serial/curve_models)
 *   - RistrettoPoint compress / decompress / elligator            (ristretto.rs == RFC 9496)
 *   - Scalar::non_adjacent_form(5), Scalar::to_radix_2w(w)        (scalar.rs)
 *   - Straus (N < 190) and Pippenger (w = 6/7/8) vartime MSM      (backend/serial/scalar_mul)
 *   - scalar arithmetic mod l (Montgomery, results canonical)     (backend/serial/u64/scalar.rs)
 * Call sites in the reference: circuit_lib.rs:187-229,363-412,491-575; util.rs:6-94; poly.rs:14-76.
 * Pinned against RFC 9496 vectors, libsodium fixtures and the Python oracle in tests/test_oracle.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint8_t u8;

/* ------------------------------------------------------------------ field: 5 x 51 bits ---- */
typedef struct { u64 v[5]; } fe;
#define M51 ((1ULL << 51) - 1)

static const fe FE_ZERO = {{0, 0, 0, 0, 0}};
static const fe FE_ONE = {{1, 0, 0, 0, 0}};
static fe FE_D, FE_D2, FE_SQRTM1, FE_INVSQRT_A_MINUS_D, FE_SQRT_AD_MINUS_ONE, FE_ONE_MINUS_D_SQ, FE_D_MINUS_ONE_SQ;

static void fe_frombytes(fe *r, const u8 *s) {
    u64 w[4];
    memcpy(w, s, 32);
    r->v[0] = w[0] & M51;
    r->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
    r->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
    r->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
    r->v[4] = (w[3] >> 12) & M51; /* bit 255 ignored, like dalek */
}
static void fe_weak_reduce(fe *r) {
    u64 c0 = r->v[0] >> 51, c1 = r->v[1] >> 51, c2 = r->v[2] >> 51, c3 = r->v[3] >> 51, c4 = r->v[4] >> 51;
    r->v[0] = (r->v[0] & M51) + c4 * 19;
    r->v[1] = (r->v[1] & M51) + c0;
    r->v[2] = (r->v[2] & M51) + c1;
    r->v[3] = (r->v[3] & M51) + c2;
    r->v[4] = (r->v[4] & M51) + c3;
}
static void fe_tobytes(u8 *s, const fe *a) {
    fe t = *a;
    fe_weak_reduce(&t);
    fe_weak_reduce(&t);
    /* now t < 2^255 + small; compute q = (t + 19) >> 255 and subtract q*p */
    u64 q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51;
    q = (t.v[2] + q) >> 51;
    q = (t.v[3] + q) >> 51;
    q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    u64 c = t.v[0] >> 51; t.v[0] &= M51;
    t.v[1] += c; c = t.v[1] >> 51; t.v[1] &= M51;
    t.v[2] += c; c = t.v[2] >> 51; t.v[2] &= M51;
    t.v[3] += c; c = t.v[3] >> 51; t.v[3] &= M51;
    t.v[4] += c; t.v[4] &= M51;
    u64 w[4];
    w[0] = t.v[0] | (t.v[1] << 51);
    w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
    w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
    w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
    memcpy(s, w, 32);
}
static void fe_add(fe *r, const fe *a, const fe *b) {
    for (int i = 0; i < 5; i++) r->v[i] = a->v[i] + b->v[i];
}
static void fe_sub(fe *r, const fe *a, const fe *b) {
    /* a + 16p - b, then one carry pass (limbs of b must be < 2^55) */
    r->v[0] = a->v[0] + 36028797018963664ULL - b->v[0];
    r->v[1] = a->v[1] + 36028797018963952ULL - b->v[1];
    r->v[2] = a->v[2] + 36028797018963952ULL - b->v[2];
    r->v[3] = a->v[3] + 36028797018963952ULL - b->v[3];
    r->v[4] = a->v[4] + 36028797018963952ULL - b->v[4];
    fe_weak_reduce(r);
}
static void fe_neg(fe *r, const fe *a) { fe_sub(r, &FE_ZERO, a); }
static void fe_mul(fe *r, const fe *a, const fe *b) {
    const u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3], a4 = a->v[4];
    const u64 b0 = b->v[0], b1 = b->v[1], b2 = b->v[2], b3 = b->v[3], b4 = b->v[4];
    const u64 b1_19 = b1 * 19, b2_19 = b2 * 19, b3_19 = b3 * 19, b4_19 = b4 * 19;
    u128 c0 = (u128)a0 * b0 + (u128)a4 * b1_19 + (u128)a3 * b2_19 + (u128)a2 * b3_19 + (u128)a1 * b4_19;
    u128 c1 = (u128)a1 * b0 + (u128)a0 * b1 + (u128)a4 * b2_19 + (u128)a3 * b3_19 + (u128)a2 * b4_19;
    u128 c2 = (u128)a2 * b0 + (u128)a1 * b1 + (u128)a0 * b2 + (u128)a4 * b3_19 + (u128)a3 * b4_19;
    u128 c3 = (u128)a3 * b0 + (u128)a2 * b1 + (u128)a1 * b2 + (u128)a0 * b3 + (u128)a4 * b4_19;
    u128 c4 = (u128)a4 * b0 + (u128)a3 * b1 + (u128)a2 * b2 + (u128)a1 * b3 + (u128)a0 * b4;
    c1 += (u64)(c0 >> 51); u64 o0 = (u64)c0 & M51;
    c2 += (u64)(c1 >> 51); u64 o1 = (u64)c1 & M51;
    c3 += (u64)(c2 >> 51); u64 o2 = (u64)c2 & M51;
    c4 += (u64)(c3 >> 51); u64 o3 = (u64)c3 & M51;
    u64 carry = (u64)(c4 >> 51); u64 o4 = (u64)c4 & M51;
    o0 += carry * 19;
    o1 += o0 >> 51; o0 &= M51;
    r->v[0] = o0; r->v[1] = o1; r->v[2] = o2; r->v[3] = o3; r->v[4] = o4;
}
static void fe_sq(fe *r, const fe *a) {
    const u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3], a4 = a->v[4];
    const u64 a3_19 = 19 * a3, a4_19 = 19 * a4;
    u128 c0 = (u128)a0 * a0 + 2 * ((u128)a1 * a4_19 + (u128)a2 * a3_19);
    u128 c1 = (u128)a3 * a3_19 + 2 * ((u128)a0 * a1 + (u128)a2 * a4_19);
    u128 c2 = (u128)a1 * a1 + 2 * ((u128)a0 * a2 + (u128)a4 * a3_19);
    u128 c3 = (u128)a4 * a4_19 + 2 * ((u128)a0 * a3 + (u128)a1 * a2);
    u128 c4 = (u128)a2 * a2 + 2 * ((u128)a0 * a4 + (u128)a1 * a3);
    c1 += (u64)(c0 >> 51); u64 o0 = (u64)c0 & M51;
    c2 += (u64)(c1 >> 51); u64 o1 = (u64)c1 & M51;
    c3 += (u64)(c2 >> 51); u64 o2 = (u64)c2 & M51;
    c4 += (u64)(c3 >> 51); u64 o3 = (u64)c3 & M51;
    u64 carry = (u64)(c4 >> 51); u64 o4 = (u64)c4 & M51;
    o0 += carry * 19;
    o1 += o0 >> 51; o0 &= M51;
    r->v[0] = o0; r->v[1] = o1; r->v[2] = o2; r->v[3] = o3; r->v[4] = o4;
}
static void fe_sqn(fe *r, const fe *a, int n) {
    fe t = *a;
    for (int i = 0; i < n; i++) fe_sq(&t, &t);
    *r = t;
}
static void fe_pow22501(fe *t250, fe *z11, const fe *z) {
    fe z2, z9, t, u, z10, z20, z50, z100;
    fe_sq(&z2, z);
    fe_sqn(&t, &z2, 2);
    fe_mul(&z9, &t, z);
    fe_mul(z11, &z9, &z2);
    fe_sq(&t, z11);
    fe_mul(&t, &t, &z9); /* 2^5-1 */
    fe_sqn(&u, &t, 5); fe_mul(&z10, &u, &t);
    fe_sqn(&u, &z10, 10); fe_mul(&z20, &u, &z10);
    fe_sqn(&u, &z20, 20); fe_mul(&t, &u, &z20);
    fe_sqn(&u, &t, 10); fe_mul(&z50, &u, &z10);
    fe_sqn(&u, &z50, 50); fe_mul(&z100, &u, &z50);
    fe_sqn(&u, &z100, 100); fe_mul(&t, &u, &z100);
    fe_sqn(&u, &t, 50); fe_mul(t250, &u, &z50);
}
static void fe_invert(fe *r, const fe *z) {
    fe t250, z11, t;
    fe_pow22501(&t250, &z11, z);
    fe_sqn(&t, &t250, 5);
    fe_mul(r, &t, &z11);
}
static void fe_pow_p58(fe *r, const fe *z) {
    fe t250, z11, t;
    fe_pow22501(&t250, &z11, z);
    fe_sqn(&t, &t250, 2);
    fe_mul(r, &t, z);
}
static int fe_is_zero(const fe *a) {
    u8 s[32];
    fe_tobytes(s, a);
    u8 o = 0;
    for (int i = 0; i < 32; i++) o |= s[i];
    return o == 0;
}
static int fe_is_neg(const fe *a) {
    u8 s[32];
    fe_tobytes(s, a);
    return s[0] & 1;
}
static int fe_eq(const fe *a, const fe *b) {
    u8 s[32], t[32];
    fe_tobytes(s, a);
    fe_tobytes(t, b);
    return memcmp(s, t, 32) == 0;
}
static void fe_abs(fe *r, const fe *a) {
    if (fe_is_neg(a)) fe_neg(r, a); else *r = *a;
}
/* RFC 9496 SQRT_RATIO_M1 == FieldElement::sqrt_ratio_i */
static int fe_sqrt_ratio_i(fe *r, const fe *u, const fe *v) {
    fe v3, v7, t, rr, check, nu, nui;
    fe_sq(&t, v); fe_mul(&v3, &t, v);
    fe_sq(&t, &v3); fe_mul(&v7, &t, v);
    fe_mul(&t, u, &v7);
    fe_pow_p58(&t, &t);
    fe_mul(&rr, u, &v3); fe_mul(&rr, &rr, &t);
    fe_sq(&t, &rr); fe_mul(&check, v, &t);
    fe_neg(&nu, u);
    fe_mul(&nui, &nu, &FE_SQRTM1);
    int correct = fe_eq(&check, u), flipped = fe_eq(&check, &nu), flipped_i = fe_eq(&check, &nui);
    if (flipped || flipped_i) fe_mul(&rr, &rr, &FE_SQRTM1);
    fe_abs(r, &rr);
    return correct || flipped;
}

/* ------------------------------------------------------------------ curve models ----------- */
typedef struct { fe X, Y, Z, T; } ge_ext;         /* EdwardsPoint */
typedef struct { fe X, Y, Z; } ge_proj;            /* ProjectivePoint */
typedef struct { fe X, Y, Z, T; } ge_compl;        /* CompletedPoint */
typedef struct { fe Yp, Ym, Z, T2d; } ge_pniels;   /* ProjectiveNielsPoint */

static void ge_identity(ge_ext *r) { r->X = FE_ZERO; r->Y = FE_ONE; r->Z = FE_ONE; r->T = FE_ZERO; }
static void ge_to_pniels(ge_pniels *r, const ge_ext *p) {
    fe_add(&r->Yp, &p->Y, &p->X);
    fe_sub(&r->Ym, &p->Y, &p->X);
    r->Z = p->Z;
    fe_mul(&r->T2d, &p->T, &FE_D2);
}
static void ge_compl_to_ext(ge_ext *r, const ge_compl *c) {
    fe_mul(&r->X, &c->X, &c->T);
    fe_mul(&r->Y, &c->Y, &c->Z);
    fe_mul(&r->Z, &c->Z, &c->T);
    fe_mul(&r->T, &c->X, &c->Y);
}
static void ge_compl_to_proj(ge_proj *r, const ge_compl *c) {
    fe_mul(&r->X, &c->X, &c->T);
    fe_mul(&r->Y, &c->Y, &c->Z);
    fe_mul(&r->Z, &c->Z, &c->T);
}
/* EdwardsPoint + ProjectiveNiels -> Completed (sign = -1 subtracts) */
static void ge_add_pn(ge_compl *r, const ge_ext *p, const ge_pniels *q, int neg) {
    fe ypx, ymx, pp, mm, tt2d, zz, zz2;
    fe_add(&ypx, &p->Y, &p->X);
    fe_sub(&ymx, &p->Y, &p->X);
    fe_mul(&pp, &ypx, neg ? &q->Ym : &q->Yp);
    fe_mul(&mm, &ymx, neg ? &q->Yp : &q->Ym);
    fe_mul(&tt2d, &p->T, &q->T2d);
    fe_mul(&zz, &p->Z, &q->Z);
    fe_add(&zz2, &zz, &zz);
    fe_sub(&r->X, &pp, &mm);
    fe_add(&r->Y, &pp, &mm);
    if (!neg) { fe_add(&r->Z, &zz2, &tt2d); fe_sub(&r->T, &zz2, &tt2d); }
    else      { fe_sub(&r->Z, &zz2, &tt2d); fe_add(&r->T, &zz2, &tt2d); }
}
static void ge_proj_double(ge_compl *r, const ge_proj *p) {
    fe xx, yy, zz2, xpy, xpy2, yypxx, yymxx;
    fe_sq(&xx, &p->X);
    fe_sq(&yy, &p->Y);
    fe_sq(&zz2, &p->Z); fe_add(&zz2, &zz2, &zz2);
    fe_add(&xpy, &p->X, &p->Y);
    fe_sq(&xpy2, &xpy);
    fe_add(&yypxx, &yy, &xx);
    fe_sub(&yymxx, &yy, &xx);
    fe_sub(&r->X, &xpy2, &yypxx);
    r->Y = yypxx;
    r->Z = yymxx;
    fe_sub(&r->T, &zz2, &yymxx);
}
static void ge_add(ge_ext *r, const ge_ext *p, const ge_ext *q) {
    ge_pniels n; ge_compl c;
    ge_to_pniels(&n, q);
    ge_add_pn(&c, p, &n, 0);
    ge_compl_to_ext(r, &c);
}
static void ge_double(ge_ext *r, const ge_ext *p) {
    ge_proj pr = {p->X, p->Y, p->Z}; ge_compl c;
    ge_proj_double(&c, &pr);
    ge_compl_to_ext(r, &c);
}
static void ge_mul_pow2(ge_ext *r, const ge_ext *p, int k) {
    ge_proj s = {p->X, p->Y, p->Z}; ge_compl c;
    for (int i = 0; i < k - 1; i++) { ge_proj_double(&c, &s); ge_compl_to_proj(&s, &c); }
    ge_proj_double(&c, &s);
    ge_compl_to_ext(r, &c);
}

/* ------------------------------------------------------------------ ristretto -------------- */
static void ristretto_compress(u8 *out, const ge_ext *p) {
    fe u1, u2, t, inv, i1, i2, zinv, den, X = p->X, Y = p->Y, s;
    fe_add(&u1, &p->Z, &p->Y); fe_sub(&t, &p->Z, &p->Y); fe_mul(&u1, &u1, &t);
    fe_mul(&u2, &p->X, &p->Y);
    fe_sq(&t, &u2); fe_mul(&t, &t, &u1);
    fe_sqrt_ratio_i(&inv, &FE_ONE, &t);
    fe_mul(&i1, &inv, &u1); fe_mul(&i2, &inv, &u2);
    fe_mul(&t, &i1, &i2); fe_mul(&zinv, &t, &p->T);
    den = i2;
    fe_mul(&t, &p->T, &zinv);
    if (fe_is_neg(&t)) {
        fe_mul(&X, &p->Y, &FE_SQRTM1);
        fe_mul(&Y, &p->X, &FE_SQRTM1);
        fe_mul(&den, &i1, &FE_INVSQRT_A_MINUS_D);
    }
    fe_mul(&t, &X, &zinv);
    if (fe_is_neg(&t)) fe_neg(&Y, &Y);
    fe_sub(&t, &p->Z, &Y);
    fe_mul(&s, &den, &t);
    fe_abs(&s, &s);
    fe_tobytes(out, &s);
}
static int ristretto_decompress(ge_ext *r, const u8 *in) {
    fe s, ss, u1, u2, u2s, v, t, inv, dx, dy;
    u8 chk[32];
    fe_frombytes(&s, in);
    fe_tobytes(chk, &s);
    if (memcmp(chk, in, 32) != 0 || (in[0] & 1)) return 0;
    fe_sq(&ss, &s);
    fe_sub(&u1, &FE_ONE, &ss);
    fe_add(&u2, &FE_ONE, &ss);
    fe_sq(&u2s, &u2);
    fe_sq(&t, &u1); fe_mul(&t, &t, &FE_D); fe_neg(&t, &t);
    fe_sub(&v, &t, &u2s);
    fe_mul(&t, &v, &u2s);
    int ok = fe_sqrt_ratio_i(&inv, &FE_ONE, &t);
    fe_mul(&dx, &inv, &u2);
    fe_mul(&t, &inv, &dx); fe_mul(&dy, &t, &v);
    fe_add(&t, &s, &s); fe_mul(&t, &t, &dx);
    fe_abs(&r->X, &t);
    fe_mul(&r->Y, &u1, &dy);
    r->Z = FE_ONE;
    fe_mul(&r->T, &r->X, &r->Y);
    if (!ok || fe_is_neg(&r->T) || fe_is_zero(&r->Y)) return 0;
    return 1;
}
static void ristretto_elligator(ge_ext *out, const fe *r0) {
    fe r, u, v, t, t2, s, sp, c, N, w0, w1, w2, w3, mone;
    fe_neg(&mone, &FE_ONE);
    fe_sq(&t, r0); fe_mul(&r, &FE_SQRTM1, &t);
    fe_add(&t, &r, &FE_ONE); fe_mul(&u, &t, &FE_ONE_MINUS_D_SQ);
    fe_mul(&t, &r, &FE_D); fe_sub(&t, &mone, &t);
    fe_add(&t2, &r, &FE_D);
    fe_mul(&v, &t, &t2);
    int was_sq = fe_sqrt_ratio_i(&s, &u, &v);
    fe_mul(&t, &s, r0); fe_abs(&t, &t); fe_neg(&sp, &t);
    c = mone;
    if (!was_sq) { s = sp; c = r; }
    fe_sub(&t, &r, &FE_ONE); fe_mul(&t, &c, &t); fe_mul(&t, &t, &FE_D_MINUS_ONE_SQ);
    fe_sub(&N, &t, &v);
    fe_mul(&t, &s, &v); fe_add(&w0, &t, &t);
    fe_mul(&w1, &N, &FE_SQRT_AD_MINUS_ONE);
    fe_sq(&t, &s);
    fe_sub(&w2, &FE_ONE, &t);
    fe_add(&w3, &FE_ONE, &t);
    fe_mul(&out->X, &w0, &w3);
    fe_mul(&out->Y, &w2, &w1);
    fe_mul(&out->Z, &w1, &w3);
    fe_mul(&out->T, &w0, &w2);
}
static void ristretto_from_uniform(ge_ext *out, const u8 *b64) {
    fe r1, r2; ge_ext p1, p2;
    fe_frombytes(&r1, b64);
    fe_frombytes(&r2, b64 + 32);
    ristretto_elligator(&p1, &r1);
    ristretto_elligator(&p2, &r2);
    ge_add(out, &p1, &p2);
}

/* ------------------------------------------------------------------ constants -------------- */
static void fe_from_hex_le_words(fe *r, const u64 w[4]) { fe_frombytes(r, (const u8 *)w); }
static int g_init_done = 0;
static void orc_init(void) {
    if (g_init_done) return;
    static const u64 D[4] = {0x75eb4dca135978a3ULL, 0x00700a4d4141d8abULL, 0x8cc740797779e898ULL, 0x52036cee2b6ffe73ULL};
    static const u64 SQRTM1[4] = {0xc4ee1b274a0ea0b0ULL, 0x2f431806ad2fe478ULL, 0x2b4d00993dfbd7a7ULL, 0x2b8324804fc1df0bULL};
    static const u64 ISAMD[4] = {0x99c8fdaa805d40eaULL, 0x9d2f16175a4172beULL, 0x16c27b91fe01d840ULL, 0x786c8905cfaffca2ULL};
    static const u64 SADM1[4] = {0x7e97f6a0497b2e1bULL, 0xaf9d8e0c1b7854bdULL, 0x0f3cfcc931f5d1fdULL, 0x376931bf2b8348acULL};
    static const u64 OMDS[4] = {0xe27c09c1945fc176ULL, 0x2c81a138cd5e350fULL, 0x9994abddbe70dfe4ULL, 0x029072a8b2b3e0d7ULL};
    static const u64 DMOS[4] = {0x31ad5aaa44ed4d20ULL, 0xd29e4a2cb01e1999ULL, 0x4cdcd32f529b4eebULL, 0x5968b37af66c2241ULL};
    fe_from_hex_le_words(&FE_D, D);
    fe_add(&FE_D2, &FE_D, &FE_D); fe_weak_reduce(&FE_D2);
    fe_from_hex_le_words(&FE_SQRTM1, SQRTM1);
    fe_from_hex_le_words(&FE_INVSQRT_A_MINUS_D, ISAMD);
    fe_from_hex_le_words(&FE_SQRT_AD_MINUS_ONE, SADM1);
    fe_from_hex_le_words(&FE_ONE_MINUS_D_SQ, OMDS);
    fe_from_hex_le_words(&FE_D_MINUS_ONE_SQ, DMOS);
    g_init_done = 1;
}

/* ------------------------------------------------------------------ scalar recodings ------- */
/* Scalar::non_adjacent_form(w): 256 signed odd digits */
static void sc_naf(int8_t naf[256], const u8 s[32], int w) {
    u64 x[5] = {0, 0, 0, 0, 0};
    memcpy(x, s, 32);
    memset(naf, 0, 256);
    const u64 width = 1ULL << w, mask = width - 1;
    int pos = 0;
    u64 carry = 0;
    while (pos < 256) {
        int idx = pos / 64, bit = pos % 64;
        u64 buf = (bit < 64 - w) ? (x[idx] >> bit) : ((x[idx] >> bit) | (x[idx + 1] << (64 - bit)));
        u64 window = carry + (buf & mask);
        if ((window & 1) == 0) { pos += 1; continue; }
        if (window < width / 2) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)((int64_t)window - (int64_t)width); }
        pos += w;
    }
}
/* Scalar::to_radix_2w(w), 4 <= w <= 8; returns digit count */
static int sc_radix_2w(int8_t digits[64], const u8 s[32], int w) {
    u64 x[4];
    memcpy(x, s, 32);
    const u64 radix = 1ULL << w, mask = radix - 1;
    int dc = (256 + w - 1) / w;
    u64 carry = 0;
    memset(digits, 0, 64);
    for (int i = 0; i < dc; i++) {
        int bit_off = i * w, idx = bit_off / 64, b = bit_off % 64;
        u64 buf;
        if (b < 64 - w || idx == 3) buf = x[idx] >> b;
        else buf = (x[idx] >> b) | (x[idx + 1] << (64 - b));
        u64 coef = carry + (buf & mask);
        carry = (coef + radix / 2) >> w;
        digits[i] = (int8_t)((int64_t)coef - (int64_t)(carry << w));
    }
    if (w == 8) { digits[dc] += (int8_t)carry; return dc + 1; }
    digits[dc - 1] += (int8_t)(carry << w);
    return dc;
}

/* ------------------------------------------------------------------ MSM -------------------- */
/* Straus vartime (straus.rs optional_multiscalar_mul): NAF(5), tables [P,3P,..,15P] */
static void msm_straus(ge_ext *out, const u8 *scalars, const ge_ext *pts, size_t n) {
    int8_t(*nafs)[256] = malloc(n * 256);
    ge_pniels(*tab)[8] = malloc(n * sizeof(ge_pniels[8]));
    for (size_t j = 0; j < n; j++) {
        sc_naf(nafs[j], scalars + 32 * j, 5);
        ge_ext a2, cur = pts[j];
        ge_double(&a2, &pts[j]);
        ge_pniels a2n; ge_to_pniels(&a2n, &a2);
        ge_to_pniels(&tab[j][0], &cur);
        for (int k = 1; k < 8; k++) {
            ge_compl c; ge_add_pn(&c, &cur, &a2n, 0); ge_compl_to_ext(&cur, &c);
            ge_to_pniels(&tab[j][k], &cur);
        }
    }
    ge_proj r = {FE_ZERO, FE_ONE, FE_ONE};
    ge_compl t; ge_ext e;
    for (int i = 255; i >= 0; i--) {
        ge_proj_double(&t, &r);
        for (size_t j = 0; j < n; j++) {
            int d = nafs[j][i];
            if (d > 0) { ge_compl_to_ext(&e, &t); ge_add_pn(&t, &e, &tab[j][d / 2], 0); }
            else if (d < 0) { ge_compl_to_ext(&e, &t); ge_add_pn(&t, &e, &tab[j][(-d) / 2], 1); }
        }
        ge_compl_to_proj(&r, &t);
    }
    /* r (projective) -> extended */
    out->X = r.X; out->Y = r.Y; out->Z = r.Z;
    {   /* T = XY/Z: rebuild through one completed conversion: (X*Z, Y*Z, Z*Z, X*Y) */
        fe_mul(&out->X, &r.X, &r.Z); fe_mul(&out->Y, &r.Y, &r.Z); fe_sq(&out->Z, &r.Z); fe_mul(&out->T, &r.X, &r.Y);
    }
    free(nafs); free(tab);
}
/* Pippenger vartime (pippenger.rs) */
static void msm_pippenger(ge_ext *out, const u8 *scalars, const ge_ext *pts, size_t n) {
    int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
    int buckets_count = (1 << w) / 2;
    int8_t(*digs)[64] = malloc(n * 64);
    ge_pniels *pn = malloc(n * sizeof(ge_pniels));
    int dc = 0;
    for (size_t j = 0; j < n; j++) { dc = sc_radix_2w(digs[j], scalars + 32 * j, w); ge_to_pniels(&pn[j], &pts[j]); }
    ge_ext *buckets = malloc(buckets_count * sizeof(ge_ext));
    ge_ext total; int have = 0;
    for (int di = dc - 1; di >= 0; di--) {
        for (int b = 0; b < buckets_count; b++) ge_identity(&buckets[b]);
        for (size_t j = 0; j < n; j++) {
            int d = digs[j][di]; ge_compl c;
            if (d > 0) { ge_add_pn(&c, &buckets[d - 1], &pn[j], 0); ge_compl_to_ext(&buckets[d - 1], &c); }
            else if (d < 0) { ge_add_pn(&c, &buckets[-d - 1], &pn[j], 1); ge_compl_to_ext(&buckets[-d - 1], &c); }
        }
        ge_ext inter = buckets[buckets_count - 1], sum = buckets[buckets_count - 1];
        for (int b = buckets_count - 2; b >= 0; b--) { ge_add(&inter, &inter, &buckets[b]); ge_add(&sum, &sum, &inter); }
        if (!have) { total = sum; have = 1; }
        else { ge_mul_pow2(&total, &total, w); ge_add(&total, &total, &sum); }
    }
    *out = total;
    free(digs); free(pn); free(buckets);
}
static void msm_dispatch(ge_ext *out, const u8 *scalars, const ge_ext *pts, size_t n) {
    if (n == 0) { ge_identity(out); return; }
    if (n < 190) msm_straus(out, scalars, pts, n); else msm_pippenger(out, scalars, pts, n);
}

/* ------------------------------------------------------------------ exported API ----------- */
/* points cross this API as raw ge_ext structs (160 B = dalek's in-memory RistrettoPoint layout) */
int orc_point_size(void) { return (int)sizeof(ge_ext); }

int orc_decompress(const u8 *in32, size_t n, u8 *out_pts) {
    orc_init();
    for (size_t i = 0; i < n; i++)
        if (!ristretto_decompress((ge_ext *)out_pts + i, in32 + 32 * i)) return -5;
    return 0;
}
void orc_compress(const u8 *pts, size_t n, u8 *out32) {
    orc_init();
    for (size_t i = 0; i < n; i++) ristretto_compress(out32 + 32 * i, (const ge_ext *)pts + i);
}
void orc_from_uniform(const u8 *bytes64, size_t n, u8 *out_pts) {
    orc_init();
    for (size_t i = 0; i < n; i++) ristretto_from_uniform((ge_ext *)out_pts + i, bytes64 + 64 * i);
}
void orc_point_add(const u8 *a, const u8 *b, u8 *out) {
    orc_init();
    ge_add((ge_ext *)out, (const ge_ext *)a, (const ge_ext *)b);
}
/* RistrettoPoint::vartime_multiscalar_mul with dalek's size dispatch; result as a raw point */
void orc_msm_vartime(const u8 *scalars, const u8 *pts, size_t n, u8 *out_pt) {
    orc_init();
    msm_dispatch((ge_ext *)out_pt, scalars, (const ge_ext *)pts, n);
}
/* which = 0 Straus, 1 Pippenger (forced), for cross-checking the two restatements */
void orc_msm_forced(int which, const u8 *scalars, const u8 *pts, size_t n, u8 *out_pt) {
    orc_init();
    if (which == 0) msm_straus((ge_ext *)out_pt, scalars, (const ge_ext *)pts, n);
    else msm_pippenger((ge_ext *)out_pt, scalars, (const ge_ext *)pts, n);
}
/* `RistrettoPoint * Scalar` (circuit_lib.rs:491): constant-time radix-16 in dalek; result identical */
void orc_scalar_mul(const u8 *scalar, const u8 *pt, u8 *out_pt) {
    orc_init();
    msm_straus((ge_ext *)out_pt, scalar, (const ge_ext *)pt, 1);
}

/* all-host-cores variant for the reported baseline: the points are split over `threads` workers,
 * each runs the dalek-dispatch MSM on its slice, partial sums are added. */
typedef struct { const u8 *sc; const ge_ext *pts; size_t n; ge_ext out; } msm_job;
static void *msm_worker(void *arg) {
    msm_job *j = arg;
    msm_dispatch(&j->out, j->sc, j->pts, j->n);
    return NULL;
}
void orc_msm_vartime_mt(const u8 *scalars, const u8 *pts, size_t n, int threads, u8 *out_pt) {
    orc_init();
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)(n ? n : 1);
    pthread_t *th = malloc(sizeof(pthread_t) * threads);
    msm_job *jobs = malloc(sizeof(msm_job) * threads);
    size_t per = (n + threads - 1) / threads, off = 0;
    int used = 0;
    for (int t = 0; t < threads && off < n; t++, used++) {
        size_t cnt = n - off < per ? n - off : per;
        jobs[t].sc = scalars + 32 * off; jobs[t].pts = (const ge_ext *)pts + off; jobs[t].n = cnt;
        pthread_create(&th[t], NULL, msm_worker, &jobs[t]);
        off += cnt;
    }
    ge_ext acc; ge_identity(&acc);
    for (int t = 0; t < used; t++) { pthread_join(th[t], NULL); ge_add(&acc, &acc, &jobs[t].out); }
    *(ge_ext *)out_pt = acc;
    free(th); free(jobs);
}
