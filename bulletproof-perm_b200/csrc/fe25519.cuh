// fe25519.cuh - arithmetic in F_p, p = 2^255 - 19, for sm_100a.
//
// Replaces curve25519-dalek-ng 4.1.1 `FieldElement51` (serial u64 backend, the
// arithmetic under every group operation the reference performs; call sites
// /root/reference/bp-perm/src/circuit_lib.rs:187-229,363-412,491-575).
//
// Representation: 8 saturated 32-bit limbs, little endian, value in [0, 2^256)
// and only congruent mod p ("lazy"): 2^256 = 38 (mod p), so carries out of limb
// 7 are folded back by multiplying with 38.  Canonical bytes are produced only
// by fe_tobytes().  The multiplier is an 8x8 schoolbook in IMAD.WIDE.U32: PTX
// mad.lo.cc/madc.hi.cc pairs that ptxas fuses into one wide multiply-add with
// carry-in/carry-out, on two interleaved accumulators ("even" and "odd" limb
// alignment) so that no carry ever has to ripple between products.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#error "fe25519.cuh is device code"
#endif

#define FE_INLINE __device__ __forceinline__
#include "mp256.cuh"

struct fe {
    uint32_t v[8];
};

// p, little-endian limbs
__device__ __constant__ const uint32_t FE_P[8] = {0xffffffedu, 0xffffffffu, 0xffffffffu, 0xffffffffu,
                                                  0xffffffffu, 0xffffffffu, 0xffffffffu, 0x7fffffffu};

FE_INLINE void fe_set0(fe &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
FE_INLINE void fe_set1(fe &r) {
    fe_set0(r);
    r.v[0] = 1;
}
FE_INLINE void fe_copy(fe &r, const fe &a) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = a.v[i];
}

// r (9 limbs: r[0..7], top) = lo[0..7] + 38 * hi[0..7]; then top folded -> 8 limbs < 2^256.
FE_INLINE void fe_reduce16(fe &r, const uint32_t *t /*16 limbs*/) {
    // even lanes: hi0,hi2,hi4,hi6 * 38 + lo[0..7]
    uint32_t e0 = t[0], e1 = t[1], e2 = t[2], e3 = t[3], e4 = t[4], e5 = t[5], e6 = t[6], e7 = t[7], e8 = 0;
    fe_mad4(e0, e1, e2, e3, e4, e5, e6, e7, e8, t[8], t[10], t[12], t[14], 38u);
    // odd lanes: hi1,hi3,hi5,hi7 * 38 at limb offset 1 (o0 is limb 1 .. o7 is limb 8)
    uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0, o4 = 0, o5 = 0, o6 = 0, o7 = 0;
    fe_mad4_nc(o0, o1, o2, o3, o4, o5, o6, o7, t[9], t[11], t[13], t[15], 38u);
    // combine: limbs 1..8
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+&r"(e1), "+&r"(e2), "+&r"(e3), "+&r"(e4), "+&r"(e5), "+&r"(e6), "+&r"(e7), "+&r"(e8)
        : "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6), "r"(o7));
    // e8 < 2^7: fold 38*e8 into limb 0 and ripple; a final carry (value wrapped past 2^256,
    // so the remainder is < 38*e8) is folded once more without further ripple.
    uint32_t f = e8 * 38u, c;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+&r"(e0), "+&r"(e1), "+&r"(e2), "+&r"(e3), "+&r"(e4), "+&r"(e5), "+&r"(e6), "+&r"(e7), "=&r"(c)
        : "r"(f));
    e0 += c * 38u;
    r.v[0] = e0; r.v[1] = e1; r.v[2] = e2; r.v[3] = e3;
    r.v[4] = e4; r.v[5] = e5; r.v[6] = e6; r.v[7] = e7;
}

// r = a * b.  64 wide multiply-adds for the product (mp_mul8, mp256.cuh) + 8 for the fold by 38.
FE_INLINE void fe_mul(fe &r, const fe &a, const fe &b) {
    uint32_t t[16];
    mp_mul8<8>(t, a.v, b.v);
    fe_reduce16(r, t);
}

// r = a^2.  28 off-diagonal products (each counted twice by one left shift of their sum) + 8 squares
// + 8 for the fold: 44 wide multiply-adds instead of 72.  Same even/odd accumulator layout as fe_mul;
// every chain's carry-out lands on a limb that so far holds only earlier carries, so it cannot wrap.
FE_INLINE void fe_sqr(fe &r, const fe &a) {
    uint32_t ev[16], od[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { ev[i] = 0; od[i] = 0; }
    const uint32_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4], a5 = a.v[5], a6 = a.v[6],
                   a7 = a.v[7];
    fe_mad4(od[0], od[1], od[2], od[3], od[4], od[5], od[6], od[7], od[8], a1, a3, a5, a7, a0);
    fe_mad3(ev[2], ev[3], ev[4], ev[5], ev[6], ev[7], ev[8], a2, a4, a6, a0);
    fe_mad3(od[2], od[3], od[4], od[5], od[6], od[7], od[8], a2, a4, a6, a1);
    fe_mad3(ev[4], ev[5], ev[6], ev[7], ev[8], ev[9], ev[10], a3, a5, a7, a1);
    fe_mad3(od[4], od[5], od[6], od[7], od[8], od[9], od[10], a3, a5, a7, a2);
    fe_mad2(ev[6], ev[7], ev[8], ev[9], ev[10], a4, a6, a2);
    fe_mad2(od[6], od[7], od[8], od[9], od[10], a4, a6, a3);
    fe_mad2(ev[8], ev[9], ev[10], ev[11], ev[12], a5, a7, a3);
    fe_mad2(od[8], od[9], od[10], od[11], od[12], a5, a7, a4);
    fe_mad1(ev[10], ev[11], ev[12], a6, a4);
    fe_mad1(od[10], od[11], od[12], a6, a5);
    fe_mad1(ev[12], ev[13], ev[14], a7, a5);
    fe_mad1(od[12], od[13], od[14], a7, a6);
    // s = ev + (od << 32): the off-diagonal sum, < 2^511
    uint32_t s[16];
    s[0] = ev[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;"
        : "=&r"(s[1]), "=&r"(s[2]), "=&r"(s[3]), "=&r"(s[4]), "=&r"(s[5]), "=&r"(s[6]), "=&r"(s[7]), "=&r"(s[8]),
          "=&r"(s[9]), "=&r"(s[10]), "=&r"(s[11]), "=&r"(s[12]), "=&r"(s[13]), "=&r"(s[14]), "=&r"(s[15])
        : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]),
          "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]),
          "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]),
          "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
    // t = 2 s + sum a_i^2 2^(64 i)
    uint32_t t[16];
#pragma unroll
    for (int k = 15; k >= 1; k--) t[k] = __funnelshift_l(s[k - 1], s[k], 1);
    t[0] = s[0] << 1;
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+&r"(t[0]), "+&r"(t[1]), "+&r"(t[2]), "+&r"(t[3]), "+&r"(t[4]), "+&r"(t[5]), "+&r"(t[6]), "+&r"(t[7]),
          "+&r"(t[8]), "+&r"(t[9]), "+&r"(t[10]), "+&r"(t[11]), "+&r"(t[12]), "+&r"(t[13]), "+&r"(t[14]), "+&r"(t[15])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7));
    fe_reduce16(r, t);
}

// r = a + b  (inputs < 2^256, output < 2^256)
FE_INLINE void fe_add(fe &r, const fe &a, const fe &b) {
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7]), "=&r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
          "r"(b.v[6]), "r"(b.v[7]));
    // fold the carry: +38.  The ripple past limb 0 is rare -> out of line.
    uint32_t f = c * 38u;
    uint32_t old = r.v[0];
    r.v[0] = old + f;
    if (r.v[0] < old) {  // carry out of limb 0 (needs r.v[0] >= 2^32-38)
        uint32_t c2;
        asm("add.cc.u32 %0, %0, 1;\n\t"
            "addc.cc.u32 %1, %1, 0;\n\t"
            "addc.cc.u32 %2, %2, 0;\n\t"
            "addc.cc.u32 %3, %3, 0;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\t"
            "addc.cc.u32 %5, %5, 0;\n\t"
            "addc.cc.u32 %6, %6, 0;\n\t"
            "addc.u32 %7, 0, 0;"
            : "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]), "+&r"(r.v[7]),
              "=&r"(c2));
        r.v[0] += c2 * 38u;  // second wrap leaves limbs 1..7 zero and limb 0 < 38: cannot carry
    }
}

// r = a - b  (inputs < 2^256, output < 2^256)
FE_INLINE void fe_sub(fe &r, const fe &a, const fe &b) {
    uint32_t bw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7]), "=&r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
          "r"(b.v[6]), "r"(b.v[7]));
    // bw is 0 or 0xffffffff.  A borrow means the true value is r - 2^256 = r - 38 (mod p).
    uint32_t f = bw & 38u;
    uint32_t old = r.v[0];
    r.v[0] = old - f;
    if (old < f) {  // borrow out of limb 0 (rare)
        uint32_t b2;
        asm("sub.cc.u32 %0, %0, 1;\n\t"
            "subc.cc.u32 %1, %1, 0;\n\t"
            "subc.cc.u32 %2, %2, 0;\n\t"
            "subc.cc.u32 %3, %3, 0;\n\t"
            "subc.cc.u32 %4, %4, 0;\n\t"
            "subc.cc.u32 %5, %5, 0;\n\t"
            "subc.cc.u32 %6, %6, 0;\n\t"
            "subc.u32 %7, 0, 0;"
            : "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]), "+&r"(r.v[7]),
              "=&r"(b2));
        r.v[0] -= b2 & 38u;  // second wrap leaves the value >= 2^256-76: cannot borrow
    }
}

FE_INLINE void fe_neg(fe &r, const fe &a) {
    fe z;
    fe_set0(z);
    fe_sub(r, z, a);
}

// r = 2a
FE_INLINE void fe_dbl(fe &r, const fe &a) { fe_add(r, a, a); }

// Fully reduce to the canonical representative in [0, p).
FE_INLINE void fe_canon(fe &r, const fe &a) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = a.v[i];
    // two rounds of folding bit 255 (weight 19) bring the value below 2^255
#pragma unroll
    for (int round = 0; round < 2; round++) {
        uint32_t top = v[7] >> 31;
        v[7] &= 0x7fffffffu;
        uint32_t f = top * 19u;
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, 0;\n\t"
            "addc.cc.u32 %2, %2, 0;\n\t"
            "addc.cc.u32 %3, %3, 0;\n\t"
            "addc.cc.u32 %4, %4, 0;\n\t"
            "addc.cc.u32 %5, %5, 0;\n\t"
            "addc.cc.u32 %6, %6, 0;\n\t"
            "addc.u32 %7, %7, 0;"
            : "+&r"(v[0]), "+&r"(v[1]), "+&r"(v[2]), "+&r"(v[3]), "+&r"(v[4]), "+&r"(v[5]), "+&r"(v[6]), "+&r"(v[7])
            : "r"(f));
    }
    // now v < 2^255; subtract p iff v >= p, i.e. iff v + 19 has bit 255 set
    uint32_t w[8];
    asm("add.cc.u32 %0, %8, 19;\n\t"
        "addc.cc.u32 %1, %9, 0;\n\t"
        "addc.cc.u32 %2, %10, 0;\n\t"
        "addc.cc.u32 %3, %11, 0;\n\t"
        "addc.cc.u32 %4, %12, 0;\n\t"
        "addc.cc.u32 %5, %13, 0;\n\t"
        "addc.cc.u32 %6, %14, 0;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=&r"(w[0]), "=&r"(w[1]), "=&r"(w[2]), "=&r"(w[3]), "=&r"(w[4]), "=&r"(w[5]), "=&r"(w[6]), "=&r"(w[7])
        : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    bool ge = (w[7] >> 31) != 0;
    w[7] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = ge ? w[i] : v[i];
}

// predicates on the canonical value
FE_INLINE bool fe_is_zero(const fe &a) {
    fe c;
    fe_canon(c, a);
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= c.v[i];
    return o == 0;
}
FE_INLINE bool fe_is_negative(const fe &a) {  // low bit of the canonical encoding
    fe c;
    fe_canon(c, a);
    return (c.v[0] & 1u) != 0;
}
FE_INLINE bool fe_eq(const fe &a, const fe &b) {
    fe d;
    fe_sub(d, a, b);
    return fe_is_zero(d);
}
FE_INLINE void fe_cneg(fe &r, const fe &a, bool neg) {
    fe n;
    fe_neg(n, a);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = neg ? n.v[i] : a.v[i];
}
FE_INLINE void fe_abs(fe &r, const fe &a) { fe_cneg(r, a, fe_is_negative(a)); }

FE_INLINE void fe_frombytes(fe &r, const uint8_t *s) {  // 32 LE bytes, all 256 bits kept
#pragma unroll
    for (int i = 0; i < 8; i++)
        r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) |
                 ((uint32_t)s[4 * i + 3] << 24);
}
FE_INLINE void fe_tobytes(uint8_t *s, const fe &a) {
    fe c;
    fe_canon(c, a);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s[4 * i] = (uint8_t)c.v[i];
        s[4 * i + 1] = (uint8_t)(c.v[i] >> 8);
        s[4 * i + 2] = (uint8_t)(c.v[i] >> 16);
        s[4 * i + 3] = (uint8_t)(c.v[i] >> 24);
    }
}

// n squarings, rolled (keeps code size small; the loop body is one fe_mul)
static __device__ __noinline__ void fe_sqr_n(fe &r, const fe &a, int n) {
    fe t;
    fe_copy(t, a);
#pragma unroll 1
    for (int i = 0; i < n; i++) fe_sqr(t, t);
    fe_copy(r, t);
}
static __device__ __noinline__ void fe_mul_noinline(fe &r, const fe &a, const fe &b) { fe_mul(r, a, b); }

// z^(2^250-1) and z^11, the shared prefix of inversion and pow((p-5)/8).
static __device__ __noinline__ void fe_pow_2_250_1(fe &t250, fe &z11, const fe &z) {
    fe z2, z9, t, u;
    fe_sqr_n(z2, z, 1);                      // 2
    fe_sqr_n(t, z2, 2);                      // 8
    fe_mul_noinline(z9, t, z);               // 9
    fe_mul_noinline(z11, z9, z2);            // 11
    fe_sqr_n(t, z11, 1);                     // 22
    fe_mul_noinline(t, t, z9);               // 31 = 2^5-1
    fe_sqr_n(u, t, 5);
    fe_mul_noinline(t, u, t);                // 2^10-1
    fe z10;
    fe_copy(z10, t);
    fe_sqr_n(u, t, 10);
    fe_mul_noinline(t, u, z10);              // 2^20-1
    fe z20;
    fe_copy(z20, t);
    fe_sqr_n(u, t, 20);
    fe_mul_noinline(t, u, z20);              // 2^40-1
    fe_sqr_n(u, t, 10);
    fe_mul_noinline(t, u, z10);              // 2^50-1
    fe z50;
    fe_copy(z50, t);
    fe_sqr_n(u, t, 50);
    fe_mul_noinline(t, u, z50);              // 2^100-1
    fe z100;
    fe_copy(z100, t);
    fe_sqr_n(u, t, 100);
    fe_mul_noinline(t, u, z100);             // 2^200-1
    fe_sqr_n(u, t, 50);
    fe_mul_noinline(t250, u, z50);           // 2^250-1
}

// r = z^(p-2) = 1/z (0 -> 0)
static __device__ __noinline__ void fe_invert(fe &r, const fe &z) {
    fe t250, z11, t;
    fe_pow_2_250_1(t250, z11, z);
    fe_sqr_n(t, t250, 5);                    // 2^255-2^5
    fe_mul_noinline(r, t, z11);              // 2^255-21
}
// r = z^((p-5)/8) = z^(2^252-3)
static __device__ __noinline__ void fe_pow_p58(fe &r, const fe &z) {
    fe t250, z11, t;
    fe_pow_2_250_1(t250, z11, z);
    fe_sqr_n(t, t250, 2);                    // 2^252-4
    fe_mul_noinline(r, t, z);                // 2^252-3
}
