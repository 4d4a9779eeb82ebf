/* abi_smoke.c - the C ABI seen from a C11 compiler (test infrastructure).
 *
 * Proves three things the ctypes bindings cannot: (1) include/bpperm.h parses as C (gcc -std=c11 -pedantic -Werror),
 * (2) every entry point this harness uses is assigned to a pointer of the header's own type (__typeof__(&fn)), so a
 * signature that drifts from the header no longer compiles, (3) a plain C host can drive the path: dlopen the
 * library, bpp_init, one bpp_msm_vartime_host and one bpp_acproof_prove_batch / bpp_acproof_verify_batch on inputs
 * written by the test (tests/test_c_harness.py), results written back for comparison with the oracle.
 *
 * usage: abi_smoke <libbpperm_cuda.so> symbols
 *        abi_smoke <libbpperm_cuda.so> run <input file> <output file>
 * input file (little endian): u32 n_msm | n_msm x 32 scalars | n_msm x 32 points |
 *        u32 n, Q, m, mode, count, nnz[4] | wire[] | constraint[] | coeff[] x 32 | c_vec Q x 32 | g | h | G ng x 32 | H ng x 32 |
 *        u32 ng | a_L, a_R, a_O count x n x 32 | gamma count x m x 32 | seeds count x 32 | V count x m x 32
 * output file: 32 (MSM) | count x proof_len proofs | count accept bytes
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bpperm.h"

/* the POSIX idiom for dlsym into a function pointer; the pointer has the type the header declares for `name` */
#define LOAD(name)                                                          \
    __typeof__(&name) p_##name = NULL;                                      \
    *(void **)(&p_##name) = dlsym(lib, #name);                              \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

static uint8_t *take(uint8_t **cur, size_t n) { uint8_t *r = *cur; *cur += n; return r; }
static uint32_t take_u32(uint8_t **cur) { uint32_t v; memcpy(&v, *cur, 4); *cur += 4; return v; }

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s lib symbols | run in out\n", argv[0]); return 64; }
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    LOAD(bpp_init) LOAD(bpp_free) LOAD(bpp_strerror) LOAD(bpp_last_error) LOAD(bpp_launch_count)
    LOAD(bpp_msm_vartime_host) LOAD(bpp_msm_vartime) LOAD(bpp_points_upload) LOAD(bpp_points_free) LOAD(bpp_points_precompute)
    LOAD(bpp_msm_vartime_batch) LOAD(bpp_circuit_create) LOAD(bpp_circuit_create_shuffle) LOAD(bpp_circuit_free)
    LOAD(bpp_gens_create) LOAD(bpp_gens_free) LOAD(bpp_acproof_proof_len_mode) LOAD(bpp_acproof_wire_len)
    LOAD(bpp_acproof_prove_batch) LOAD(bpp_acproof_verify_batch) LOAD(bpp_acp_batch_create) LOAD(bpp_acp_batch_free)
    LOAD(bpp_acp_batch_gen_shuffle_witness) LOAD(bpp_acp_batch_commit) LOAD(bpp_acp_batch_upload_commitments)
    LOAD(bpp_acp_batch_prove) LOAD(bpp_acp_batch_verify) LOAD(bpp_acp_batch_gather_accept) LOAD(bpp_comm_unique_id)
    LOAD(bpp_comm_init) LOAD(bpp_msm_sharded_dev) LOAD(bpp_inner_product) LOAD(bpp_scalar_invert)
    if (strcmp(argv[2], "symbols") == 0) {
        /* no device needed: argument validation and the pure byte helpers */
        if (p_bpp_acproof_proof_len_mode(104, 2) != 32 * (13 + 2 * 7)) return 3;
        if (p_bpp_acproof_proof_len_mode(104, 1) != 32 * (11 + 2 * 104)) return 3;
        if (p_bpp_acproof_wire_len(104, 2) != 1 + 32 * (13 + 2 * 7)) return 3;
        if (p_bpp_acp_batch_prove(NULL) != BPP_ERR_INVALID_ARG) return 4;
        if (p_bpp_acp_batch_verify(NULL, NULL) != BPP_ERR_INVALID_ARG) return 4;
        if (p_bpp_msm_vartime(NULL, NULL, 0, NULL, 0, 0, NULL, NULL) != BPP_ERR_INVALID_ARG) return 4;
        if (p_bpp_circuit_create_shuffle(NULL, 52, NULL) != BPP_ERR_INVALID_ARG) return 4;
        if (strcmp(p_bpp_strerror(BPP_ERR_INVALID_ARG), "invalid argument") != 0) return 5;
        printf("symbols ok\n");
        return 0;
    }
    if (argc < 5) return 64;
    FILE *f = fopen(argv[3], "rb");
    if (!f) return 66;
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *buf = malloc((size_t)len), *cur = buf;
    if (fread(buf, 1, (size_t)len, f) != (size_t)len) return 66;
    fclose(f);
    bpp_ctx *ctx = NULL;
    int rc = p_bpp_init(0, &ctx);
    if (rc) { fprintf(stderr, "bpp_init: %s\n", p_bpp_strerror(rc)); return 10; }
    /* 1. the trait-level MSM */
    uint32_t n_msm = take_u32(&cur);
    uint8_t *sc = take(&cur, 32 * (size_t)n_msm), *pts = take(&cur, 32 * (size_t)n_msm);
    uint8_t msm_out[32];
    rc = p_bpp_msm_vartime_host(ctx, sc, n_msm, BPP_FMT_COMPRESSED, pts, n_msm, msm_out);
    if (rc) { fprintf(stderr, "msm: %s (%s)\n", p_bpp_strerror(rc), p_bpp_last_error(ctx)); return 11; }
    /* 2. prove + verify a batch */
    uint32_t n = take_u32(&cur), Q = take_u32(&cur), m = take_u32(&cur), mode = take_u32(&cur), count = take_u32(&cur);
    uint32_t nnz[4];
    size_t total = 0;
    for (int i = 0; i < 4; i++) { nnz[i] = take_u32(&cur); total += nnz[i]; }
    uint32_t *wire = malloc(4 * total + 4), *cons = malloc(4 * total + 4);
    memcpy(wire, take(&cur, 4 * total), 4 * total);
    memcpy(cons, take(&cur, 4 * total), 4 * total);
    uint8_t *coeff = take(&cur, 32 * total), *cvec = take(&cur, 32 * (size_t)Q);
    uint8_t *g = take(&cur, 32), *h = take(&cur, 32);
    uint32_t ng = take_u32(&cur);
    uint8_t *G = take(&cur, 32 * (size_t)ng), *H = take(&cur, 32 * (size_t)ng);
    uint8_t *aL = take(&cur, 32 * (size_t)count * n), *aR = take(&cur, 32 * (size_t)count * n), *aO = take(&cur, 32 * (size_t)count * n);
    uint8_t *gamma = take(&cur, 32 * (size_t)count * m), *seeds = take(&cur, 32 * (size_t)count), *V = take(&cur, 32 * (size_t)count * m);
    if (cur - buf != len) { fprintf(stderr, "input length mismatch\n"); return 12; }
    bpp_circuit *cir = NULL;
    bpp_gens *gens = NULL;
    rc = p_bpp_circuit_create(ctx, n, Q, m, nnz, wire, cons, coeff, cvec, &cir);
    if (!rc) rc = p_bpp_gens_create(ctx, g, h, G, H, ng, 6, &gens);
    if (rc) { fprintf(stderr, "setup: %s (%s)\n", p_bpp_strerror(rc), p_bpp_last_error(ctx)); return 13; }
    size_t plen = p_bpp_acproof_proof_len_mode(n, (int)mode);
    uint8_t *proofs = malloc(plen * count), *accept = malloc(count);
    static const uint8_t label[4] = {'t', 'e', 's', 't'};
    rc = p_bpp_acproof_prove_batch(ctx, cir, gens, (int)mode, count, aL, aR, aO, gamma, seeds, V, label, 4, proofs);
    if (!rc) rc = p_bpp_acproof_verify_batch(ctx, cir, gens, (int)mode, count, proofs, V, label, 4, NULL, accept);
    if (rc) { fprintf(stderr, "prove/verify: %s (%s)\n", p_bpp_strerror(rc), p_bpp_last_error(ctx)); return 14; }
    if (p_bpp_launch_count(ctx) == 0) return 15;
    FILE *o = fopen(argv[4], "wb");
    if (!o) return 73;
    fwrite(msm_out, 1, 32, o);
    fwrite(proofs, 1, plen * count, o);
    fwrite(accept, 1, count, o);
    fclose(o);
    p_bpp_gens_free(ctx, gens);
    p_bpp_circuit_free(ctx, cir);
    p_bpp_free(ctx);
    printf("run ok\n");
    return 0;
}
