// capi_vec.cu - host entry points (C ABI) of the scalar-vector operators.
#include "bpperm_internal.hpp"
#include "vec_kernels.cuh"

// Stages host buffers into one device arena: [in0 | in1 | ... | out]; returns device pointers.
struct vec_stage {
    bpp_ctx *ctx;
    uint8_t *base = nullptr;
    size_t used = 0;
    int rc = BPP_OK;
    vec_stage(bpp_ctx *c, size_t total_bytes) : ctx(c) {
        if (cudaSetDevice(c->device) != cudaSuccess) { rc = BPP_ERR_CUDA; return; }
        rc = grow(c, &c->d_vec, &c->cap_vec, total_bytes + 256);
        base = c->d_vec;
    }
    uint32_t *put(const uint8_t *host, size_t bytes) {
        if (rc) return nullptr;
        uint8_t *p = base + used;
        used += (bytes + 31) & ~(size_t)31;
        if (host && cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = BPP_ERR_CUDA;
        return (uint32_t *)p;
    }
    int get(uint8_t *host, const uint32_t *dev, size_t bytes) {
        if (rc) return rc;
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); return BPP_ERR_CUDA; }
        return BPP_OK;
    }
};

// util.rs panics on length mismatches (util.rs:9-11,26-28,44-46,86-88); the ABI cannot see the two
// lengths of a Vec, so callers pass both and a mismatch is BPP_ERR_LENGTH_MISMATCH.
extern "C" int bpp_inner_product(bpp_ctx *ctx, const uint8_t *a, size_t na, const uint8_t *b, size_t nb, uint8_t out[32]) {
    if (!ctx || !out || (!a && na) || (!b && nb)) return BPP_ERR_INVALID_ARG;
    if (na != nb) return BPP_ERR_LENGTH_MISMATCH;
    if (na == 0) { memset(out, 0, 32); return BPP_OK; }
    vec_stage st(ctx, 64 * na + 64);
    uint32_t *da = st.put(a, 32 * na), *db = st.put(b, 32 * na), *dout = st.put(nullptr, 32);
    if (st.rc) return st.rc;
    k_rows_dot<<<1, 256, 0, ctx->stream>>>(da, db, (uint32_t)na, dout);
    return st.get(out, dout, 32);
}

extern "C" int bpp_hadamard_V(bpp_ctx *ctx, const uint8_t *a, size_t na, const uint8_t *b, size_t nb, uint8_t *out) {
    if (!ctx || (!out && na) || (!a && na) || (!b && nb)) return BPP_ERR_INVALID_ARG;
    if (na != nb) return BPP_ERR_LENGTH_MISMATCH;
    if (na == 0) return BPP_OK;
    vec_stage st(ctx, 96 * na);
    uint32_t *da = st.put(a, 32 * na), *db = st.put(b, 32 * na), *dout = st.put(nullptr, 32 * na);
    if (st.rc) return st.rc;
    k_hadamard<<<(unsigned)((na + 127) / 128), 128, 0, ctx->stream>>>(da, db, (uint32_t)na, dout);
    return st.get(out, dout, 32 * na);
}

// vm_mult(a, b): a has len_a entries, b is `rows` rows of `cols` entries; needs len_a == cols; out[rows]
extern "C" int bpp_vm_mult(bpp_ctx *ctx, const uint8_t *a, size_t len_a, const uint8_t *b, size_t rows, size_t cols,
                           uint8_t *out) {
    if (!ctx || !a || !b || !out || rows == 0 || cols == 0) return BPP_ERR_INVALID_ARG;
    if (len_a != cols) return BPP_ERR_LENGTH_MISMATCH;
    vec_stage st(ctx, 32 * (len_a + rows * cols + rows) + 128);
    uint32_t *da = st.put(a, 32 * len_a), *db = st.put(b, 32 * rows * cols), *dout = st.put(nullptr, 32 * rows);
    if (st.rc) return st.rc;
    k_rows_dot<<<(unsigned)rows, 256, 0, ctx->stream>>>(da, db, (uint32_t)cols, dout);
    return st.get(out, dout, 32 * rows);
}

// mv_mult(a, b): a is `rows` x `cols`, b has len_b entries; needs rows == len_b; out[cols]
extern "C" int bpp_mv_mult(bpp_ctx *ctx, const uint8_t *a, size_t rows, size_t cols, const uint8_t *b, size_t len_b,
                           uint8_t *out) {
    if (!ctx || !a || !b || !out || rows == 0 || cols == 0) return BPP_ERR_INVALID_ARG;
    if (len_b != rows) return BPP_ERR_LENGTH_MISMATCH;
    vec_stage st(ctx, 32 * (len_b + rows * cols + cols) + 128);
    uint32_t *da = st.put(a, 32 * rows * cols), *db = st.put(b, 32 * len_b), *dout = st.put(nullptr, 32 * cols);
    if (st.rc) return st.rc;
    k_cols_dot<<<(unsigned)cols, 256, 0, ctx->stream>>>(da, db, (uint32_t)rows, (uint32_t)cols, dout);
    return st.get(out, dout, 32 * cols);
}

extern "C" int bpp_exp_iter(bpp_ctx *ctx, const uint8_t x[32], size_t count, uint8_t *out) {
    if (!ctx || !x || (!out && count)) return BPP_ERR_INVALID_ARG;
    if (count == 0) return BPP_OK;
    vec_stage st(ctx, 32 * count + 64);
    uint32_t *dx = st.put(x, 32), *dout = st.put(nullptr, 32 * count);
    if (st.rc) return st.rc;
    k_exp_iter_fib<<<1, 32, 0, ctx->stream>>>(dx, (uint32_t)count, dout);
    return st.get(out, dout, 32 * count);
}

extern "C" int bpp_scalar_powers(bpp_ctx *ctx, const uint8_t x[32], size_t first, size_t count, uint8_t *out) {
    if (!ctx || !x || (!out && count) || first + count >= (1ull << 32)) return BPP_ERR_INVALID_ARG;
    if (count == 0) return BPP_OK;
    vec_stage st(ctx, 32 * count + 64);
    uint32_t *dx = st.put(x, 32), *dout = st.put(nullptr, 32 * count);
    if (st.rc) return st.rc;
    k_scalar_powers<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(dx, (uint32_t)first, (uint32_t)count, dout);
    return st.get(out, dout, 32 * count);
}

extern "C" int bpp_scalar_exp(bpp_ctx *ctx, const uint8_t x[32], uint32_t pow, uint8_t out[32]) {
    if (!ctx || !x || !out) return BPP_ERR_INVALID_ARG;
    vec_stage st(ctx, 128);
    uint32_t *dx = st.put(x, 32), *dout = st.put(nullptr, 32);
    if (st.rc) return st.rc;
    k_scalar_exp<<<1, 32, 0, ctx->stream>>>(dx, pow, dout);
    return st.get(out, dout, 32);
}

extern "C" int bpp_scalar_invert(bpp_ctx *ctx, const uint8_t *a, size_t n, uint8_t *out) {
    if (!ctx || (!a && n) || (!out && n)) return BPP_ERR_INVALID_ARG;
    if (n == 0) return BPP_OK;
    vec_stage st(ctx, 64 * n);
    uint32_t *da = st.put(a, 32 * n), *dout = st.put(nullptr, 32 * n);
    if (st.rc) return st.rc;
    k_scalar_invert<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(da, (uint32_t)n, dout);
    return st.get(out, dout, 32 * n);
}

extern "C" int bpp_scalar_from_wide(bpp_ctx *ctx, const uint8_t *in64, size_t n, uint8_t *out) {
    if (!ctx || (!in64 && n) || (!out && n)) return BPP_ERR_INVALID_ARG;
    if (n == 0) return BPP_OK;
    vec_stage st(ctx, 96 * n);
    uint32_t *da = st.put(in64, 64 * n), *dout = st.put(nullptr, 32 * n);
    if (st.rc) return st.rc;
    k_scalar_from_wide<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(da, (uint32_t)n, dout);
    return st.get(out, dout, 32 * n);
}

extern "C" int bpp_scalar_reduce(bpp_ctx *ctx, const uint8_t *in32, size_t n, uint8_t *out) {
    if (!ctx || (!in32 && n) || (!out && n)) return BPP_ERR_INVALID_ARG;
    if (n == 0) return BPP_OK;
    vec_stage st(ctx, 64 * n);
    uint32_t *da = st.put(in32, 32 * n), *dout = st.put(nullptr, 32 * n);
    if (st.rc) return st.rc;
    k_scalar_reduce<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(da, (uint32_t)n, dout);
    return st.get(out, dout, 32 * n);
}

// VecPoly3 is four vectors of n scalars, passed as one 4 x n array (poly.rs:21-27)
extern "C" int bpp_vecpoly3_special_inner_product(bpp_ctx *ctx, const uint8_t *lhs, const uint8_t *rhs, size_t n,
                                                  uint8_t out_t1_t6[192]) {
    if (!ctx || !lhs || !rhs || !out_t1_t6 || n == 0) return BPP_ERR_INVALID_ARG;
    vec_stage st(ctx, 256 * n + 1024);
    uint32_t *dl = st.put(lhs, 128 * n), *dr = st.put(rhs, 128 * n), *dd = st.put(nullptr, 9 * 32), *dout = st.put(nullptr, 192);
    if (st.rc) return st.rc;
    k_vecpoly3_nine_dots<<<9, 256, 0, ctx->stream>>>(dl, dr, (uint32_t)n, dd);
    ctx->launches++;
    k_vecpoly3_combine<<<1, 32, 0, ctx->stream>>>(dd, dout);
    return st.get(out_t1_t6, dout, 192);
}

extern "C" int bpp_vecpoly3_eval(bpp_ctx *ctx, const uint8_t *coeffs, size_t n, const uint8_t x[32], uint8_t *out) {
    if (!ctx || !coeffs || !x || !out || n == 0) return BPP_ERR_INVALID_ARG;
    vec_stage st(ctx, 160 * n + 128);
    uint32_t *dc = st.put(coeffs, 128 * n), *dx = st.put(x, 32), *dout = st.put(nullptr, 32 * n);
    if (st.rc) return st.rc;
    k_vecpoly3_eval<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(dc, dx, (uint32_t)n, dout);
    return st.get(out, dout, 32 * n);
}

extern "C" int bpp_poly6_eval(bpp_ctx *ctx, const uint8_t t1_t6[192], const uint8_t x[32], uint8_t out[32]) {
    if (!ctx || !t1_t6 || !x || !out) return BPP_ERR_INVALID_ARG;
    vec_stage st(ctx, 512);
    uint32_t *dt = st.put(t1_t6, 192), *dx = st.put(x, 32), *dout = st.put(nullptr, 32);
    if (st.rc) return st.rc;
    k_poly6_eval<<<1, 32, 0, ctx->stream>>>(dt, dx, dout);
    return st.get(out, dout, 32);
}
