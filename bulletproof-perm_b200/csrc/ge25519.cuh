// ge25519.cuh - twisted Edwards (a = -1) / Ristretto255 point arithmetic for sm_100a.
//
// Replaces curve25519-dalek-ng 4.1.1 `EdwardsPoint` / `RistrettoPoint`
// (`ProjectiveNielsPoint`/`AffineNielsPoint` mixed adds, `ProjectivePoint::double`,
// `RistrettoPoint::compress`, `CompressedRistretto::decompress`, `ct_eq`) as used at
// /root/reference/bp-perm/src/circuit_lib.rs:187-229 (MSM), :231-233,368-412 (compress),
// :532 (decompress), :541 (equality).  Formulas: add-2008-hwcd-3 / dbl-2008-hwcd;
// encodings: RFC 9496 sections 4.3.1-4.3.3.
#pragma once
// Bounds checks of the debug variant (tools/build_variants.sh dbg "-DBPP_DEBUG_BOUNDS"; compute-sanitizer is closed on
// this pool): every index computed from device data is asserted against the extent the host allocated before it is
// used.  A violated assertion traps and the next CUDA call of the library reports it.  Off in the product build.
#ifdef BPP_DEBUG_BOUNDS
#include <assert.h>
#define BPP_ASSERT(cond) assert(cond)
#else
#define BPP_ASSERT(cond) ((void)0)
#endif
#include "fe25519.cuh"

struct ge_ext {  // extended coordinates: x = X/Z, y = Y/Z, T = XY/Z
    fe X, Y, Z, T;
};
struct ge_niels {  // affine Niels form of (x, y): (y+x, y-x, 2d*x*y); identity = (1, 1, 0)
    fe yp, ym, t2d;
};

// curve constants, little-endian 32-bit limbs
__device__ __constant__ const uint32_t GE_D[8] = {0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du,
                                                  0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu};
__device__ __constant__ const uint32_t GE_D2[8] = {0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au,
                                                   0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu};
__device__ __constant__ const uint32_t GE_SQRTM1[8] = {0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u,
                                                       0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u};
__device__ __constant__ const uint32_t GE_INVSQRT_A_MINUS_D[8] = {0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u,
                                                                  0xfe01d840u, 0x16c27b91u, 0xcfaffca2u, 0x786c8905u};

__device__ __constant__ const uint32_t GE_SQRT_AD_MINUS_ONE[8] = {0x497b2e1bu, 0x7e97f6a0u, 0x1b7854bdu, 0xaf9d8e0cu,
                                                               0x31f5d1fdu, 0x0f3cfcc9u, 0x2b8348acu, 0x376931bfu};
__device__ __constant__ const uint32_t GE_ONE_MINUS_D_SQ[8] = {0x945fc176u, 0xe27c09c1u, 0xcd5e350fu, 0x2c81a138u,
                                                            0xbe70dfe4u, 0x9994abddu, 0xb2b3e0d7u, 0x029072a8u};
__device__ __constant__ const uint32_t GE_D_MINUS_ONE_SQ[8] = {0x44ed4d20u, 0x31ad5aaau, 0xb01e1999u, 0xd29e4a2cu,
                                                            0x529b4eebu, 0x4cdcd32fu, 0xf66c2241u, 0x5968b37au};

FE_INLINE void fe_const(fe &r, const uint32_t *c) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c[i];
}

FE_INLINE void ge_identity(ge_ext &r) {
    fe_set0(r.X);
    fe_set1(r.Y);
    fe_set1(r.Z);
    fe_set0(r.T);
}

// r = p + q (q affine Niels), or p - q when neg.  7 field multiplications.
FE_INLINE void ge_madd(ge_ext &r, const ge_ext &p, const ge_niels &q, bool neg) {
    fe a, b, c, d, e, f, g, h, qp, qm, qt;
#pragma unroll
    for (int i = 0; i < 8; i++) {  // -q = (ym, yp, -t2d)
        qp.v[i] = neg ? q.ym.v[i] : q.yp.v[i];
        qm.v[i] = neg ? q.yp.v[i] : q.ym.v[i];
    }
    fe_sub(a, p.Y, p.X);
    fe_mul(a, a, qm);
    fe_add(b, p.Y, p.X);
    fe_mul(b, b, qp);
    fe_mul(c, p.T, q.t2d);
    fe_dbl(d, p.Z);
    fe_sub(e, b, a);
    fe_add(h, b, a);
    fe_sub(f, d, c);
    fe_add(g, d, c);
#pragma unroll
    for (int i = 0; i < 8; i++) {  // negating q swaps the roles of f and g
        qt.v[i] = neg ? g.v[i] : f.v[i];
        g.v[i] = neg ? f.v[i] : g.v[i];
        f.v[i] = qt.v[i];
    }
    fe_mul(r.X, e, f);
    fe_mul(r.Y, g, h);
    fe_mul(r.Z, f, g);
    fe_mul(r.T, e, h);
}

// r = p + q, both extended.  9 field multiplications.
FE_INLINE void ge_add(ge_ext &r, const ge_ext &p, const ge_ext &q) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(a, p.Y, p.X);
    fe_sub(t, q.Y, q.X);
    fe_mul(a, a, t);
    fe_add(b, p.Y, p.X);
    fe_add(t, q.Y, q.X);
    fe_mul(b, b, t);
    fe_const(t, GE_D2);
    fe_mul(c, p.T, q.T);
    fe_mul(c, c, t);
    fe_mul(d, p.Z, q.Z);
    fe_dbl(d, d);
    fe_sub(e, b, a);
    fe_add(h, b, a);
    fe_sub(f, d, c);
    fe_add(g, d, c);
    fe_mul(r.X, e, f);
    fe_mul(r.Y, g, h);
    fe_mul(r.Z, f, g);
    fe_mul(r.T, e, h);
}

// r = 2p.  4 squarings + 4 multiplications.
FE_INLINE void ge_double(ge_ext &r, const ge_ext &p) {
    fe a, b, c, e, f, g, h, t;
    fe_sqr(a, p.X);
    fe_sqr(b, p.Y);
    fe_sqr(c, p.Z);
    fe_dbl(c, c);
    fe_add(h, a, b);
    fe_add(t, p.X, p.Y);
    fe_sqr(t, t);
    fe_sub(e, h, t);
    fe_sub(g, a, b);
    fe_add(f, c, g);
    fe_mul(r.X, e, f);
    fe_mul(r.Y, g, h);
    fe_mul(r.Z, f, g);
    fe_mul(r.T, e, h);
}

// r = 2p without the T coordinate (r.T is left unspecified): for runs of doublings, where only the
// last one before an addition needs T.  4 squarings + 3 multiplications.
FE_INLINE void ge_double_p2(ge_ext &r, const ge_ext &p) {
    fe a, b, c, e, f, g, h, t;
    fe_sqr(a, p.X);
    fe_sqr(b, p.Y);
    fe_sqr(c, p.Z);
    fe_dbl(c, c);
    fe_add(h, a, b);
    fe_add(t, p.X, p.Y);
    fe_sqr(t, t);
    fe_sub(e, h, t);
    fe_sub(g, a, b);
    fe_add(f, c, g);
    fe_mul(r.X, e, f);
    fe_mul(r.Y, g, h);
    fe_mul(r.Z, f, g);
}

static __device__ __noinline__ void ge_add_noinline(ge_ext &r, const ge_ext &p, const ge_ext &q) { ge_add(r, p, q); }
static __device__ __noinline__ void ge_double_noinline(ge_ext &r, const ge_ext &p) { ge_double(r, p); }

FE_INLINE void ge_neg(ge_ext &r, const ge_ext &p) {
    fe_neg(r.X, p.X);
    fe_copy(r.Y, p.Y);
    fe_copy(r.Z, p.Z);
    fe_neg(r.T, p.T);
}

// affine (x, y) -> Niels
FE_INLINE void ge_affine_to_niels(ge_niels &r, const fe &x, const fe &y) {
    fe t, d2;
    fe_add(r.yp, y, x);
    fe_sub(r.ym, y, x);
    fe_mul(t, x, y);
    fe_const(d2, GE_D2);
    fe_mul(r.t2d, t, d2);
}

// RFC 9496 4.2: (was_square, r) with r = |sqrt(u/v)| or |sqrt(i*u/v)|
static __device__ __noinline__ bool fe_sqrt_ratio_i(fe &r, const fe &u, const fe &v) {
    fe v3, v7, t, rr, check, i, nu, nui;
    fe_mul_noinline(t, v, v);
    fe_mul_noinline(v3, t, v);
    fe_mul_noinline(t, v3, v3);
    fe_mul_noinline(v7, t, v);
    fe_mul_noinline(t, u, v7);
    fe_pow_p58(t, t);
    fe_mul_noinline(rr, u, v3);
    fe_mul_noinline(rr, rr, t);
    fe_mul_noinline(t, rr, rr);
    fe_mul_noinline(check, v, t);
    fe_const(i, GE_SQRTM1);
    fe_neg(nu, u);
    fe_mul_noinline(nui, nu, i);
    bool correct = fe_eq(check, u);
    bool flipped = fe_eq(check, nu);
    bool flipped_i = fe_eq(check, nui);
    if (flipped || flipped_i) fe_mul_noinline(rr, rr, i);
    fe_abs(r, rr);
    return correct || flipped;
}

// RFC 9496 4.3.2 Encode
static __device__ __noinline__ void ge_compress(uint8_t *out32, const ge_ext &p) {
    fe u1, u2, t, inv, i1, i2, zinv, den, X, Y, one, c;
    fe_add(u1, p.Z, p.Y);
    fe_sub(t, p.Z, p.Y);
    fe_mul_noinline(u1, u1, t);
    fe_mul_noinline(u2, p.X, p.Y);
    fe_mul_noinline(t, u2, u2);
    fe_mul_noinline(t, t, u1);
    fe_set1(one);
    fe_sqrt_ratio_i(inv, one, t);
    fe_mul_noinline(i1, inv, u1);
    fe_mul_noinline(i2, inv, u2);
    fe_mul_noinline(t, i1, i2);
    fe_mul_noinline(zinv, t, p.T);
    fe_copy(X, p.X);
    fe_copy(Y, p.Y);
    fe_copy(den, i2);
    fe_mul_noinline(t, p.T, zinv);
    if (fe_is_negative(t)) {
        fe_const(c, GE_SQRTM1);
        fe_mul_noinline(X, p.Y, c);
        fe_mul_noinline(Y, p.X, c);
        fe_const(c, GE_INVSQRT_A_MINUS_D);
        fe_mul_noinline(den, i1, c);
    }
    fe_mul_noinline(t, X, zinv);
    if (fe_is_negative(t)) fe_neg(Y, Y);
    fe_sub(t, p.Z, Y);
    fe_mul_noinline(t, den, t);
    fe_abs(t, t);
    fe_tobytes(out32, t);
}

// RFC 9496 4.3.1 Decode -> affine (x, y); returns false for invalid encodings
static __device__ __noinline__ bool ge_decompress(fe &x, fe &y, const uint8_t *in32) {
    fe s, c, ss, u1, u2, u2s, v, t, inv, dx, dy, one, d;
    fe_frombytes(s, in32);
    fe_canon(c, s);
    bool canonical = true;
#pragma unroll
    for (int i = 0; i < 8; i++) canonical = canonical && (c.v[i] == s.v[i]);
    bool ok = canonical && ((s.v[0] & 1u) == 0);
    fe_set1(one);
    fe_mul_noinline(ss, s, s);
    fe_sub(u1, one, ss);
    fe_add(u2, one, ss);
    fe_mul_noinline(u2s, u2, u2);
    fe_const(d, GE_D);
    fe_mul_noinline(t, u1, u1);
    fe_mul_noinline(t, t, d);
    fe_neg(t, t);
    fe_sub(v, t, u2s);
    fe_mul_noinline(t, v, u2s);
    bool was_sq = fe_sqrt_ratio_i(inv, one, t);
    fe_mul_noinline(dx, inv, u2);
    fe_mul_noinline(t, inv, dx);
    fe_mul_noinline(dy, t, v);
    fe_add(t, s, s);
    fe_mul_noinline(t, t, dx);
    fe_abs(x, t);
    fe_mul_noinline(y, u1, dy);
    fe_mul_noinline(t, x, y);
    ok = ok && was_sq && !fe_is_negative(t) && !fe_is_zero(y);
    return ok;
}

// RFC 9496 4.3.4 MAP (dalek RistrettoPoint::elligator_ristretto_flavor)
static __device__ __noinline__ void ge_elligator(ge_ext &out, const fe &r0) {
    fe i, d, one, r, u, v, t, t2, s, sp, c, N, w0, w1, w2, w3, k;
    fe_const(i, GE_SQRTM1);
    fe_const(d, GE_D);
    fe_set1(one);
    fe_mul_noinline(t, r0, r0);
    fe_mul_noinline(r, i, t);
    fe_add(t, r, one);
    fe_const(k, GE_ONE_MINUS_D_SQ);
    fe_mul_noinline(u, t, k);
    fe_mul_noinline(t, r, d);
    fe_neg(t2, one);
    fe_sub(t, t2, t);   // -1 - r*d
    fe_add(t2, r, d);
    fe_mul_noinline(v, t, t2);
    bool was_sq = fe_sqrt_ratio_i(s, u, v);
    fe_mul_noinline(t, s, r0);
    fe_abs(t, t);
    fe_neg(sp, t);
    fe_neg(c, one);
    if (!was_sq) {
        fe_copy(s, sp);
        fe_copy(c, r);
    }
    fe_sub(t, r, one);
    fe_mul_noinline(t, c, t);
    fe_const(k, GE_D_MINUS_ONE_SQ);
    fe_mul_noinline(t, t, k);
    fe_sub(N, t, v);
    fe_mul_noinline(t, s, v);
    fe_dbl(w0, t);
    fe_const(k, GE_SQRT_AD_MINUS_ONE);
    fe_mul_noinline(w1, N, k);
    fe_mul_noinline(t, s, s);
    fe_sub(w2, one, t);
    fe_add(w3, one, t);
    fe_mul_noinline(out.X, w0, w3);
    fe_mul_noinline(out.Y, w2, w1);
    fe_mul_noinline(out.Z, w1, w3);
    fe_mul_noinline(out.T, w0, w2);
}

// dalek RistrettoPoint::from_uniform_bytes (RFC 9496 one-way map): 64 bytes -> point
static __device__ __noinline__ void ge_from_uniform(ge_ext &out, const uint8_t *b64) {
    fe r1, r2;
    fe_frombytes(r1, b64);
    fe_frombytes(r2, b64 + 32);
    r1.v[7] &= 0x7fffffffu;
    r2.v[7] &= 0x7fffffffu;
    ge_ext p1, p2;
    ge_elligator(p1, r1);
    ge_elligator(p2, r2);
    ge_add_noinline(out, p1, p2);
}

// Ristretto equality: X1*Y2 == Y1*X2 or X1*X2 == Y1*Y2
static __device__ __noinline__ bool ge_ristretto_eq(const ge_ext &p, const ge_ext &q) {
    fe a, b;
    fe_mul_noinline(a, p.X, q.Y);
    fe_mul_noinline(b, p.Y, q.X);
    bool e1 = fe_eq(a, b);
    fe_mul_noinline(a, p.X, q.X);
    fe_mul_noinline(b, p.Y, q.Y);
    return e1 || fe_eq(a, b);
}

// ---- memory layout helpers: a ge_ext is 32 u32 = 128 B, a ge_niels is 24 u32 = 96 B ----
FE_INLINE void fe_load(fe &r, const uint32_t *p) {  // 32-byte aligned
    uint4 lo = *reinterpret_cast<const uint4 *>(p);
    uint4 hi = *reinterpret_cast<const uint4 *>(p + 4);
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
}
FE_INLINE void fe_load_nc(fe &r, const uint32_t *p) {  // read-only path
    uint4 lo = __ldg(reinterpret_cast<const uint4 *>(p));
    uint4 hi = __ldg(reinterpret_cast<const uint4 *>(p + 4));
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
}
FE_INLINE void fe_store(uint32_t *p, const fe &a) {
    *reinterpret_cast<uint4 *>(p) = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    *reinterpret_cast<uint4 *>(p + 4) = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
FE_INLINE void ge_load(ge_ext &r, const uint32_t *p) {
    fe_load(r.X, p);
    fe_load(r.Y, p + 8);
    fe_load(r.Z, p + 16);
    fe_load(r.T, p + 24);
}
FE_INLINE void ge_store(uint32_t *p, const ge_ext &a) {
    fe_store(p, a.X);
    fe_store(p + 8, a.Y);
    fe_store(p + 16, a.Z);
    fe_store(p + 24, a.T);
}
FE_INLINE void ge_niels_load(ge_niels &r, const uint32_t *p) {
    fe_load_nc(r.yp, p);
    fe_load_nc(r.ym, p + 8);
    fe_load_nc(r.t2d, p + 16);
}
FE_INLINE void ge_niels_store(uint32_t *p, const ge_niels &a) {
    fe_store(p, a.yp);
    fe_store(p + 8, a.ym);
    fe_store(p + 16, a.t2d);
}
