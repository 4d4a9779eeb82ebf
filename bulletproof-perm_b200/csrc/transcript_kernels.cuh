// transcript_kernels.cuh - the Fiat-Shamir steps of the batched shuffle prover / verifier on the
// device: one thread per proof runs that proof's Merlin transcript (merlin_dev.cuh) and writes the
// challenge scalars straight into the proof's scalar block.  Labels and order follow
// /root/reference/bp-perm/src/circuit_lib.rs:231-233 (A_I, A_O, S), :133-138 (y, z), :368-412
// (T1, T3, T4 - carrying T_3 in `reference` mode, :391 -, T5, T6), :425-432 (x); the `fixed` mode tail
// follows bulletproofs 4.0.0 r1cs/prover + inner_product_proof.rs (t_x, t_x_blinding, e_blinding,
// w, "ipp v1" domain separator, L/R -> u per round).
#pragma once
#include "acproof_kernels.cuh"
#include "merlin_dev.cuh"

#define TR_THREADS 32

// proto: the transcript after Transcript::new(label) + arithmetic_domain_sep(n), identical for every
// proof (hashed once on the host).  pts8: B x 8 x 32 compressed (A_I, A_O, S, T1, T3, T4, T5, T6).
__global__ void __launch_bounds__(TR_THREADS) k_tr_prove_yz(const uint64_t *__restrict__ proto,
                                                            const uint8_t *__restrict__ pts8, acp_layout lay,
                                                            uint32_t B, uint64_t *__restrict__ states,
                                                            uint32_t *__restrict__ blk) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    merlin_tr t;
    t.load(proto);
    const uint8_t *pt = pts8 + 256 * (size_t)p;
    t.append_message(MERLIN_LABEL("A_I"), pt, 32);
    t.append_message(MERLIN_LABEL("A_O"), pt + 32, 32);
    t.append_message(MERLIN_LABEL("S"), pt + 64, 32);
    sc c;
    t.challenge_scalar(MERLIN_LABEL("y"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.y), c);
    t.challenge_scalar(MERLIN_LABEL("z"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.z), c);
    t.store(states + MERLIN_STATE_WORDS * (size_t)p);
}

__global__ void __launch_bounds__(TR_THREADS) k_tr_prove_x(const uint8_t *__restrict__ pts8, acp_layout lay, uint32_t B,
                                                           int mode, uint64_t *__restrict__ states,
                                                           uint32_t *__restrict__ blk) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    merlin_tr t;
    t.load(states + MERLIN_STATE_WORDS * (size_t)p);
    const uint8_t *pt = pts8 + 256 * (size_t)p + 96;
    t.append_message(MERLIN_LABEL("T1"), pt, 32);
    t.append_message(MERLIN_LABEL("T3"), pt + 32, 32);
    t.append_message(MERLIN_LABEL("T4"), mode == 0 ? pt + 32 : pt + 64, 32);  // circuit_lib.rs:391 appends T_3 under "T4"
    t.append_message(MERLIN_LABEL("T5"), pt + 96, 32);
    t.append_message(MERLIN_LABEL("T6"), pt + 128, 32);
    sc c;
    t.challenge_scalar(MERLIN_LABEL("x"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.x), c);
    t.store(states + MERLIN_STATE_WORDS * (size_t)p);
}

FE_INLINE void tr_scalar_bytes(uint8_t out[32], const uint32_t *src) {
    for (int i = 0; i < 8; i++) {
        uint32_t v = src[i];
        out[4 * i] = (uint8_t)v; out[4 * i + 1] = (uint8_t)(v >> 8); out[4 * i + 2] = (uint8_t)(v >> 16);
        out[4 * i + 3] = (uint8_t)(v >> 24);
    }
}

// `fixed` mode: t_x, t_x_blinding, e_blinding (contiguous at lay.that) -> w; inner-product domain separator
__global__ void __launch_bounds__(TR_THREADS) k_tr_prove_w(acp_layout lay, uint32_t B, uint64_t *__restrict__ states,
                                                           uint32_t *__restrict__ blk) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    merlin_tr t;
    t.load(states + MERLIN_STATE_WORDS * (size_t)p);
    uint8_t s[32];
    tr_scalar_bytes(s, ACP_PTR(blk, lay, p, lay.that));
    t.append_message(MERLIN_LABEL("t_x"), s, 32);
    tr_scalar_bytes(s, ACP_PTR(blk, lay, p, lay.that + 1));
    t.append_message(MERLIN_LABEL("t_x_blinding"), s, 32);
    tr_scalar_bytes(s, ACP_PTR(blk, lay, p, lay.that + 2));
    t.append_message(MERLIN_LABEL("e_blinding"), s, 32);
    sc c;
    t.challenge_scalar(MERLIN_LABEL("w"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.wq), c);
    t.append_message(MERLIN_LABEL("dom-sep"), (const uint8_t *)"ipp v1", 6);
    t.append_u64(MERLIN_LABEL("n"), lay.np);
    t.store(states + MERLIN_STATE_WORDS * (size_t)p);
}

// `fixed` mode, round j: L_j, R_j (B x 2 lg x 32 compressed) -> u_j
__global__ void __launch_bounds__(TR_THREADS) k_tr_prove_u(const uint8_t *__restrict__ lr, acp_layout lay, uint32_t B,
                                                           uint32_t j, uint64_t *__restrict__ states,
                                                           uint32_t *__restrict__ blk) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    merlin_tr t;
    t.load(states + MERLIN_STATE_WORDS * (size_t)p);
    const uint8_t *q = lr + 64 * ((size_t)lay.lg * p + j);
    t.append_message(MERLIN_LABEL("L"), q, 32);
    t.append_message(MERLIN_LABEL("R"), q + 32, 32);
    sc c;
    t.challenge_scalar(MERLIN_LABEL("u"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.u + j), c);
    t.store(states + MERLIN_STATE_WORDS * (size_t)p);
}

// Verifier: replays the whole transcript of one proof.  tx3 (B x 3 x 32: t_x, t_x_blinding, e_blinding)
// and lr are used in `fixed` mode only (lay.lg > 0).
__global__ void __launch_bounds__(TR_THREADS) k_tr_verify(const uint64_t *__restrict__ proto,
                                                          const uint8_t *__restrict__ pts8,
                                                          const uint8_t *__restrict__ tx3,
                                                          const uint8_t *__restrict__ lr, acp_layout lay, uint32_t B,
                                                          int mode, uint32_t *__restrict__ blk) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    merlin_tr t;
    t.load(proto);
    const uint8_t *pt = pts8 + 256 * (size_t)p;
    t.append_message(MERLIN_LABEL("A_I"), pt, 32);
    t.append_message(MERLIN_LABEL("A_O"), pt + 32, 32);
    t.append_message(MERLIN_LABEL("S"), pt + 64, 32);
    sc c;
    t.challenge_scalar(MERLIN_LABEL("y"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.y), c);
    t.challenge_scalar(MERLIN_LABEL("z"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.z), c);
    t.append_message(MERLIN_LABEL("T1"), pt + 96, 32);
    t.append_message(MERLIN_LABEL("T3"), pt + 128, 32);
    t.append_message(MERLIN_LABEL("T4"), mode == 0 ? pt + 128 : pt + 160, 32);
    t.append_message(MERLIN_LABEL("T5"), pt + 192, 32);
    t.append_message(MERLIN_LABEL("T6"), pt + 224, 32);
    t.challenge_scalar(MERLIN_LABEL("x"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.x), c);
    if (mode != 2) return;
    const uint8_t *s3 = tx3 + 96 * (size_t)p;
    t.append_message(MERLIN_LABEL("t_x"), s3, 32);
    t.append_message(MERLIN_LABEL("t_x_blinding"), s3 + 32, 32);
    t.append_message(MERLIN_LABEL("e_blinding"), s3 + 64, 32);
    t.challenge_scalar(MERLIN_LABEL("w"), c);
    sc_store(ACP_PTR(blk, lay, p, lay.wq), c);
    t.append_message(MERLIN_LABEL("dom-sep"), (const uint8_t *)"ipp v1", 6);
    t.append_u64(MERLIN_LABEL("n"), lay.np);
    for (uint32_t j = 0; j < lay.lg; j++) {   // an identity encoding is rejected in k_acp_decompress_lr
        const uint8_t *q = lr + 64 * ((size_t)lay.lg * p + j);
        t.append_message(MERLIN_LABEL("L"), q, 32);
        t.append_message(MERLIN_LABEL("R"), q + 32, 32);
        t.challenge_scalar(MERLIN_LABEL("u"), c);
        sc_store(ACP_PTR(blk, lay, p, lay.u + j), c);
    }
}

// Scripted transcript for tests (one thread): script = sequence of records
//   op (1 B: 0 = append_message, 1 = challenge_bytes) | label_len (1 B) | label | n (4 B LE) | msg (op 0 only)
// starting from Transcript::new(first record's message); challenge outputs are concatenated into out.
__global__ void k_tr_script(const uint8_t *__restrict__ script, uint32_t len, uint8_t *__restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    merlin_tr t;
    for (int i = 0; i < 25; i++) t.st[i] = 0;
    // Strobe128::new("Merlin v1.0")
    const uint8_t init[18] = {1, 168, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
    for (uint32_t i = 0; i < 18; i++) t.xor_byte(i, init[i]);
    keccak_f1600_dev(t.st);
    t.pos = 0; t.pos_begin = 0; t.cur_flags = 0;
    t.meta_ad((const uint8_t *)"Merlin v1.0", 11, false);
    uint32_t o = 0, w = 0;
    while (o < len) {
        uint32_t op = script[o], ll = script[o + 1];
        const char *label = (const char *)(script + o + 2);
        o += 2 + ll;
        uint32_t n = (uint32_t)script[o] | ((uint32_t)script[o + 1] << 8) | ((uint32_t)script[o + 2] << 16) |
                     ((uint32_t)script[o + 3] << 24);
        o += 4;
        if (op == 0) {
            t.append_message(label, ll, script + o, n);
            o += n;
        } else {
            t.challenge_bytes(label, ll, out + w, n);
            w += n;
        }
    }
}
