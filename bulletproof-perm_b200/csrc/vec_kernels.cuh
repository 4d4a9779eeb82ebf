// vec_kernels.cuh - scalar-vector operators mod l on sm_100a: the reference's util.rs / poly.rs.
//
//   inner_product   util.rs:84-94        hadamard_V  util.rs:6-20      vm_mult  util.rs:22-38
//   mv_mult         util.rs:40-56        exp_iter    util.rs:63-65,139-157 (Fibonacci exponents, as coded)
//   scalar_exp      util.rs:67-82        Scalar::invert x n  circuit_lib.rs:273-275
//   VecPoly3::special_inner_product  poly.rs:39-55   VecPoly3::eval  poly.rs:57-76   Poly6::eval  poly.rs:14-18
//
// Vectors are arrays of 32-byte little-endian scalars (8 x u32).  Products are reduced per element
// and sums are block-reduced with exact modular adds, so the bytes do not depend on the reduction
// order.  These kernels are HBM/latency bound (32 B per element streams); the sizes of the protocol
// (n = 104 .. 16384) never fill a B200 - batching across proofs does (acproof_kernels.cuh).
#pragma once
#include "sc25519.cuh"

SC_INLINE void sc_shfl_down(sc &r, const sc &a, int d) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], d);
}

// Sum over the block; valid in thread 0.  blockDim.x must be a multiple of 32 (<= 896: 28 warps x 9 words of `sh`).
// The summands are reduced (< l < 2^253), so up to 1024 of them add up exactly in nine 32-bit limbs: the shuffle
// steps are plain 9-limb additions (one carry chain, no conditional subtraction of l per step) and the total is
// reduced mod l ONCE, by thread 0 - ~130 instructions per thread instead of ~600 with a modular addition per step
// (the block reduction was two thirds of the batched dot-product kernels' instructions).  Integer addition is
// associative, so the bytes still do not depend on the reduction order.
SC_INLINE void sc_add9(uint32_t a[9], const uint32_t o[9]) {
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, %8, %17;"
        : "+&r"(a[0]), "+&r"(a[1]), "+&r"(a[2]), "+&r"(a[3]), "+&r"(a[4]), "+&r"(a[5]), "+&r"(a[6]), "+&r"(a[7]), "+&r"(a[8])
        : "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]));
}
// Integer sum over the warp in nine limbs (not reduced); valid in lane 0.
SC_INLINE void warp_sum9(uint32_t a[9], const sc &mine) {
    uint32_t o[9];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = mine.v[i];
    a[8] = 0;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
        for (int i = 0; i < 9; i++) o[i] = __shfl_down_sync(0xffffffffu, a[i], d);
        sc_add9(a, o);
    }
}
// nine-limb integer -> canonical scalar
SC_INLINE void sc_from9(sc &r, const uint32_t a[9]) {
    uint32_t wide[16];
#pragma unroll
    for (int i = 0; i < 9; i++) wide[i] = a[i];
#pragma unroll
    for (int i = 9; i < 16; i++) wide[i] = 0;
    sc_from_wide(r, wide);
}
SC_INLINE void block_sum_sc(sc &total, const sc &mine, uint32_t *sh /* 32 x 8 u32 */) {
    uint32_t a[9], o[9];
    warp_sum9(a, mine);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();   // `sh` may still be read from a previous call
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; i++) sh[9 * wid + i] = a[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (uint32_t w = 1; w < nw; w++) {
#pragma unroll
            for (int i = 0; i < 9; i++) o[i] = sh[9 * w + i];
            sc_add9(a, o);
        }
        sc_from9(total, a);
    } else {
        sc_set0(total);
    }
}

// Montgomery-form sum of products: sum a_i*b_i/R; thread-local accumulation then block reduce.
SC_INLINE void dot_partial(sc &acc, const uint32_t *__restrict__ a, size_t sa, const uint32_t *__restrict__ b,
                           size_t sb, uint32_t n) {
    sc x, y, p;
    sc_set0(acc);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        sc_load(x, a + 8 * (size_t)i * sa);
        sc_load(y, b + 8 * (size_t)i * sb);
        sc_mont(p, x, y);
        sc_add(acc, acc, p);
    }
}

// out[row] = <a, b[row]> for `rows` rows of length n (vm_mult); rows == 1 is inner_product.
// Inputs canonical; one block per row.
static __global__ void __launch_bounds__(256) k_rows_dot(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                                  uint32_t n, uint32_t *__restrict__ out) {
    __shared__ __align__(16) uint32_t sh[32 * 8];
    const uint32_t row = blockIdx.x;
    sc acc, tot, k;
    dot_partial(acc, a, 1, b + 8 * (size_t)row * n, 1, n);
    block_sum_sc(tot, acc, sh);
    if (threadIdx.x == 0) {
        sc_const(k, SC_R2);
        sc_mont(tot, tot, k);  // leave Montgomery form
        sc_store(out + 8 * (size_t)row, tot);
    }
}

// out[col] = sum_i a[i][col] * b[i]  (mv_mult: a is rows x cols, row-major); one block per column.
static __global__ void __launch_bounds__(256) k_cols_dot(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                                  uint32_t rows, uint32_t cols, uint32_t *__restrict__ out) {
    __shared__ __align__(16) uint32_t sh[32 * 8];
    const uint32_t col = blockIdx.x;
    sc acc, tot, k;
    dot_partial(acc, a + 8 * (size_t)col, cols, b, 1, rows);
    block_sum_sc(tot, acc, sh);
    if (threadIdx.x == 0) {
        sc_const(k, SC_R2);
        sc_mont(tot, tot, k);
        sc_store(out + 8 * (size_t)col, tot);
    }
}

static __global__ void k_hadamard(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t n,
                           uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sc x, y, r;
    sc_load(x, a + 8 * (size_t)i);
    sc_load(y, b + 8 * (size_t)i);
    sc_mul(r, x, y);
    sc_store(out + 8 * (size_t)i, r);
}

// util.rs exp_iter as coded: state (x = 1, next = x0); each step returns next, then next *= x; x = returned.
// A serial chain by construction (x^F(i) = x^F(i-1) * x^F(i-2)); one thread.
static __global__ void k_exp_iter_fib(const uint32_t *__restrict__ x0, uint32_t count, uint32_t *__restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    sc base, nxt, ret;
    sc_const(base, SC_R);  // 1 in Montgomery form
    sc_load(nxt, x0);
    sc_to_mont(nxt, nxt);
#pragma unroll 1
    for (uint32_t i = 0; i < count; i++) {
        ret = nxt;
        sc_mont_noinline(nxt, nxt, base);
        base = ret;
        sc s;
        sc_from_mont(s, ret);
        sc_store(out + 8 * (size_t)i, s);
    }
}

// out[i] = x^(i + first) by square-and-multiply, one thread per power (standard powers 1, x, x^2, ...)
static __global__ void k_scalar_powers(const uint32_t *__restrict__ x0, uint32_t first, uint32_t count,
                                uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    sc xm, acc;
    sc_load(xm, x0);
    sc_to_mont(xm, xm);
    sc_const(acc, SC_R);
    uint32_t e = i + first;
#pragma unroll 1
    for (int bit = 31; bit >= 0; bit--) {
        sc_mont_noinline(acc, acc, acc);
        if ((e >> bit) & 1u) sc_mont_noinline(acc, acc, xm);
    }
    sc_from_mont(acc, acc);
    sc_store(out + 8 * (size_t)i, acc);
}

static __global__ void k_scalar_invert(const uint32_t *__restrict__ a, uint32_t n, uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sc x, r;
    sc_load(x, a + 8 * (size_t)i);
    sc_invert(r, x);
    sc_store(out + 8 * (size_t)i, r);
}

static __global__ void k_scalar_reduce(const uint32_t *__restrict__ a, uint32_t n, uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sc x;
    sc_load(x, a + 8 * (size_t)i);
    sc_reduce256(x, x);
    sc_store(out + 8 * (size_t)i, x);
}

static __global__ void k_scalar_from_wide(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = in[16 * (size_t)i + k];
    sc r;
    sc_from_wide(r, w);
    sc_store(out + 8 * (size_t)i, r);
}

// VecPoly3::eval: out[i] = c0[i] + x*(c1[i] + x*(c2[i] + x*c3[i])); c is 4 x n
static __global__ void k_vecpoly3_eval(const uint32_t *__restrict__ c, const uint32_t *__restrict__ x0, uint32_t n,
                                uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sc x, c0, c1, c2, c3, t;
    sc_load(x, x0);
    sc_load(c0, c + 8 * ((size_t)0 * n + i));
    sc_load(c1, c + 8 * ((size_t)1 * n + i));
    sc_load(c2, c + 8 * ((size_t)2 * n + i));
    sc_load(c3, c + 8 * ((size_t)3 * n + i));
    sc_mul(t, x, c3);
    sc_add(t, t, c2);
    sc_mul(t, x, t);
    sc_add(t, t, c1);
    sc_mul(t, x, t);
    sc_add(t, t, c0);
    sc_store(out + 8 * (size_t)i, t);
}

// VecPoly3::special_inner_product: the 9 inner products -> t1..t6 (poly.rs:39-55).  One block per
// inner product (blockIdx.x = 0..8), then block 0's thread 0 of a second launch combines - here the
// combination is done by a tiny second kernel to keep this one race-free.
static __global__ void __launch_bounds__(256) k_vecpoly3_nine_dots(const uint32_t *__restrict__ l, const uint32_t *__restrict__ r,
                                                            uint32_t n, uint32_t *__restrict__ dots /* 9 x 8 */) {
    __shared__ __align__(16) uint32_t sh[32 * 8];
    // (lhs index, rhs index) per poly.rs:40-45
    const int LI[9] = {1, 1, 2, 2, 3, 1, 3, 2, 3};
    const int RI[9] = {0, 1, 0, 1, 0, 3, 1, 3, 3};
    const int k = blockIdx.x;
    sc acc, tot, r2;
    dot_partial(acc, l + 8 * (size_t)LI[k] * n, 1, r + 8 * (size_t)RI[k] * n, 1, n);
    block_sum_sc(tot, acc, sh);
    if (threadIdx.x == 0) {
        sc_const(r2, SC_R2);
        sc_mont(tot, tot, r2);
        sc_store(dots + 8 * k, tot);
    }
}
static __global__ void k_vecpoly3_combine(const uint32_t *__restrict__ dots, uint32_t *__restrict__ t6) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    sc d[9], t;
#pragma unroll
    for (int k = 0; k < 9; k++) sc_load(d[k], dots + 8 * k);
    sc_store(t6 + 0, d[0]);           // t1 = <l1,r0>
    sc_add(t, d[1], d[2]);            // t2 = <l1,r1> + <l2,r0>
    sc_store(t6 + 8, t);
    sc_add(t, d[3], d[4]);            // t3 = <l2,r1> + <l3,r0>
    sc_store(t6 + 16, t);
    sc_add(t, d[5], d[6]);            // t4 = <l1,r3> + <l3,r1>
    sc_store(t6 + 24, t);
    sc_store(t6 + 32, d[7]);          // t5 = <l2,r3>
    sc_store(t6 + 40, d[8]);          // t6 = <l3,r3>
}

// Poly6::eval (poly.rs:14-18): x*(t1 + x*(t2 + ... + x*t6)); scalar_exp(x, pow) (util.rs:75-82)
static __global__ void k_poly6_eval(const uint32_t *__restrict__ t6, const uint32_t *__restrict__ x0, uint32_t *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    sc x, acc, c;
    sc_load(x, x0);
    sc_load(acc, t6 + 40);
#pragma unroll 1
    for (int k = 4; k >= 0; k--) {
        sc_mul_noinline(acc, x, acc);
        sc_load(c, t6 + 8 * k);
        sc_add(acc, acc, c);
    }
    sc_mul_noinline(acc, x, acc);
    sc_store(out, acc);
}
static __global__ void k_scalar_exp(const uint32_t *__restrict__ x0, uint32_t pow, uint32_t *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    sc x, acc;
    sc_load(x, x0);
    sc_set_u32(acc, 1);
#pragma unroll 1
    for (uint32_t i = 0; i < pow; i++) sc_mul_noinline(acc, acc, x);
    sc_store(out, acc);
}
