// mp256.cuh - multi-limb multiply building blocks shared by the field (fe25519.cuh) and scalar
// (sc25519.cuh) arithmetic: carry-chained IMAD.WIDE.U32 rows on two interleaved accumulators.
//
// PTX mad.lo.cc/madc.hi.cc pairs are fused by ptxas into one wide multiply-add with carry in/out.  A row
// multiplies every other limb of `a` by one limb of `b`, so consecutive products land on disjoint 64-bit
// lanes and one carry chain covers the row; rows of even limb alignment accumulate in `ev` (ev[k] sits at
// limb k), rows of odd alignment in `od` (od[k] sits at limb k + 1): no carry ever ripples between
// products, and the two accumulators are added once at the end.
// All multi-instruction asm blocks mark their outputs early-clobber: an input whose VALUE equals an
// in/out operand's initial value could otherwise share its register and be overwritten before it is read.
#pragma once
#include <stdint.h>

#ifndef FE_INLINE
#define FE_INLINE __device__ __forceinline__
#endif

// c[0..7] += (a0,a1,a2,a3) * b placed on 64-bit lanes (c0c1),(c2c3),(c4c5),(c6c7);
// the carry out of limb 7 is added into c8.
FE_INLINE void fe_mad4(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t &c4, uint32_t &c5,
                       uint32_t &c6, uint32_t &c7, uint32_t &c8, uint32_t a0, uint32_t a1, uint32_t a2,
                       uint32_t a3, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2), "+&r"(c3), "+&r"(c4), "+&r"(c5), "+&r"(c6), "+&r"(c7), "+&r"(c8)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// same, when the carry out of limb 7 is known to be zero
FE_INLINE void fe_mad4_nc(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t &c4, uint32_t &c5,
                          uint32_t &c6, uint32_t &c7, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                          uint32_t b) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2), "+&r"(c3), "+&r"(c4), "+&r"(c5), "+&r"(c6), "+&r"(c7)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// shorter carry chains of the same form: c[0..2k) += (x0..x(k-1)) * y, carry out added into ct
FE_INLINE void fe_mad3(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t &c4, uint32_t &c5,
                       uint32_t &ct, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2), "+&r"(c3), "+&r"(c4), "+&r"(c5), "+&r"(ct)
        : "r"(x0), "r"(x1), "r"(x2), "r"(y));
}
FE_INLINE void fe_mad2(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t &ct, uint32_t x0,
                       uint32_t x1, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2), "+&r"(c3), "+&r"(ct)
        : "r"(x0), "r"(x1), "r"(y));
}
FE_INLINE void fe_mad1(uint32_t &c0, uint32_t &c1, uint32_t &ct, uint32_t x0, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+&r"(c0), "+&r"(c1), "+&r"(ct)
        : "r"(x0), "r"(y));
}

// t[0 .. 8+NB) = a[0..8) * b[0..NB)   (NB = 8: full 256 x 256; NB = 4: 256 x 128).  8*NB wide multiply-adds.
template <int NB>
FE_INLINE void mp_mul8(uint32_t *t, const uint32_t *a, const uint32_t *b) {
    uint32_t ev[16], od[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { ev[i] = 0; od[i] = 0; }
#pragma unroll
    for (int i = 0; i < NB; i += 2) {
        // b[i], i even: a0,a2,a4,a6 land on even limbs i+j; a1,a3,a5,a7 on odd limbs
        fe_mad4(ev[i], ev[i + 1], ev[i + 2], ev[i + 3], ev[i + 4], ev[i + 5], ev[i + 6], ev[i + 7], ev[i + 8],
                a[0], a[2], a[4], a[6], b[i]);
        fe_mad4(od[i], od[i + 1], od[i + 2], od[i + 3], od[i + 4], od[i + 5], od[i + 6], od[i + 7], od[i + 8],
                a[1], a[3], a[5], a[7], b[i]);
        // b[i+1], odd: a1,a3,a5,a7 land on even limbs (i+1)+j; a0,a2,a4,a6 on odd limbs
        if (i + 1 < NB - 1) {
            fe_mad4(ev[i + 2], ev[i + 3], ev[i + 4], ev[i + 5], ev[i + 6], ev[i + 7], ev[i + 8], ev[i + 9],
                    ev[i + 10], a[1], a[3], a[5], a[7], b[i + 1]);
        } else {   // last row: the product fits 8 + NB limbs, no carry out
            fe_mad4_nc(ev[i + 2], ev[i + 3], ev[i + 4], ev[i + 5], ev[i + 6], ev[i + 7], ev[i + 8], ev[i + 9],
                       a[1], a[3], a[5], a[7], b[i + 1]);
        }
        fe_mad4(od[i], od[i + 1], od[i + 2], od[i + 3], od[i + 4], od[i + 5], od[i + 6], od[i + 7], od[i + 8],
                a[0], a[2], a[4], a[6], b[i + 1]);
    }
    // t = ev + (od << 32)
    t[0] = ev[0];
    if (NB == 8) {
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
            : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8]),
              "=&r"(t[9]), "=&r"(t[10]), "=&r"(t[11]), "=&r"(t[12]), "=&r"(t[13]), "=&r"(t[14]), "=&r"(t[15])
            : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]),
              "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]),
              "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]),
              "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
    } else {
        asm("add.cc.u32 %0, %11, %22;\n\t"
            "addc.cc.u32 %1, %12, %23;\n\t"
            "addc.cc.u32 %2, %13, %24;\n\t"
            "addc.cc.u32 %3, %14, %25;\n\t"
            "addc.cc.u32 %4, %15, %26;\n\t"
            "addc.cc.u32 %5, %16, %27;\n\t"
            "addc.cc.u32 %6, %17, %28;\n\t"
            "addc.cc.u32 %7, %18, %29;\n\t"
            "addc.cc.u32 %8, %19, %30;\n\t"
            "addc.cc.u32 %9, %20, %31;\n\t"
            "addc.u32 %10, %21, %32;"
            : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8]),
              "=&r"(t[9]), "=&r"(t[10]), "=&r"(t[11])
            : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]),
              "r"(ev[9]), "r"(ev[10]), "r"(ev[11]),
              "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]),
              "r"(od[8]), "r"(od[9]), "r"(od[10]));
    }
}

// q[0..8) = low 256 bits of a[0..8) * b[0..8): only the 36 products with i + j <= 7 (carries and high
// halves that would land on limb 8 and above go to scratch limbs and are dropped).
FE_INLINE void mp_mul8_lo(uint32_t *q, const uint32_t *a, const uint32_t *b) {
    uint32_t ev[10], od[10];
#pragma unroll
    for (int i = 0; i < 10; i++) { ev[i] = 0; od[i] = 0; }
    // b0: a0,a2,a4,a6 -> limbs 0,2,4,6;  a1,a3,a5,a7 -> limbs 1,3,5,7
    fe_mad4(ev[0], ev[1], ev[2], ev[3], ev[4], ev[5], ev[6], ev[7], ev[8], a[0], a[2], a[4], a[6], b[0]);
    fe_mad4(od[0], od[1], od[2], od[3], od[4], od[5], od[6], od[7], od[8], a[1], a[3], a[5], a[7], b[0]);
    // b1: a1,a3,a5 -> limbs 2,4,6;  a0,a2,a4,a6 -> limbs 1,3,5,7
    fe_mad3(ev[2], ev[3], ev[4], ev[5], ev[6], ev[7], ev[8], a[1], a[3], a[5], b[1]);
    fe_mad4(od[0], od[1], od[2], od[3], od[4], od[5], od[6], od[7], od[8], a[0], a[2], a[4], a[6], b[1]);
    // b2: a0,a2,a4 -> limbs 2,4,6;  a1,a3,a5 -> limbs 3,5,7
    fe_mad3(ev[2], ev[3], ev[4], ev[5], ev[6], ev[7], ev[8], a[0], a[2], a[4], b[2]);
    fe_mad3(od[2], od[3], od[4], od[5], od[6], od[7], od[8], a[1], a[3], a[5], b[2]);
    // b3: a1,a3 -> limbs 4,6;  a0,a2,a4 -> limbs 3,5,7
    fe_mad2(ev[4], ev[5], ev[6], ev[7], ev[8], a[1], a[3], b[3]);
    fe_mad3(od[2], od[3], od[4], od[5], od[6], od[7], od[8], a[0], a[2], a[4], b[3]);
    // b4: a0,a2 -> limbs 4,6;  a1,a3 -> limbs 5,7
    fe_mad2(ev[4], ev[5], ev[6], ev[7], ev[8], a[0], a[2], b[4]);
    fe_mad2(od[4], od[5], od[6], od[7], od[8], a[1], a[3], b[4]);
    // b5: a1 -> limb 6;  a0,a2 -> limbs 5,7
    fe_mad1(ev[6], ev[7], ev[8], a[1], b[5]);
    fe_mad2(od[4], od[5], od[6], od[7], od[8], a[0], a[2], b[5]);
    // b6: a0 -> limb 6;  a1 -> limb 7
    fe_mad1(ev[6], ev[7], ev[8], a[0], b[6]);
    fe_mad1(od[6], od[7], od[8], a[1], b[6]);
    // b7: a0 -> limb 7
    fe_mad1(od[6], od[7], od[8], a[0], b[7]);
    q[0] = ev[0];
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=&r"(q[1]), "=&r"(q[2]), "=&r"(q[3]), "=&r"(q[4]), "=&r"(q[5]), "=&r"(q[6]), "=&r"(q[7])
        : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
          "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]));
}
