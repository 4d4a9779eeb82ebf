// merlin_dev.cuh - Merlin 3.0.0 transcripts (STROBE-128 over Keccak-f[1600]) on the device, one
// thread per transcript.
//
// Replaces, for batches of independent proofs, merlin::Transcript + the reference's
// TranscriptProtocol extension trait (/root/reference/bp-perm/src/transcript_protocol.rs:12-68:
// arithmetic_domain_sep :27-30, append_scalar :32-34, append_point :45-47, challenge_scalar :62-67).
// A single transcript is a serial sponge (host_merlin.hpp keeps that form for single proofs and for
// the common prefix); thousands of independent transcripts are data parallel, and keeping them on
// the device removes every host round trip between the protocol's kernels.
// Same labels, same framing, same bytes as the host class: tests drive both with the same scripts.
#pragma once
#include <stdint.h>

#include "sc25519.cuh"

#define MERLIN_STATE_WORDS 26  // 25 Keccak lanes + one word holding pos | pos_begin << 8 | cur_flags << 16

__device__ __constant__ const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

__device__ __forceinline__ uint64_t keccak_rol(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }

// Keccak-f[1600], lanes in registers, fully unrolled rounds' inner structure (rho/pi written out)
static __device__ __noinline__ void keccak_f1600_dev(uint64_t *st) {
    uint64_t a00 = st[0], a01 = st[1], a02 = st[2], a03 = st[3], a04 = st[4];
    uint64_t a05 = st[5], a06 = st[6], a07 = st[7], a08 = st[8], a09 = st[9];
    uint64_t a10 = st[10], a11 = st[11], a12 = st[12], a13 = st[13], a14 = st[14];
    uint64_t a15 = st[15], a16 = st[16], a17 = st[17], a18 = st[18], a19 = st[19];
    uint64_t a20 = st[20], a21 = st[21], a22 = st[22], a23 = st[23], a24 = st[24];
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21;
        uint64_t c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22, c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23;
        uint64_t c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
        uint64_t d0 = c4 ^ keccak_rol(c1, 1), d1 = c0 ^ keccak_rol(c2, 1), d2 = c1 ^ keccak_rol(c3, 1);
        uint64_t d3 = c2 ^ keccak_rol(c4, 1), d4 = c3 ^ keccak_rol(c0, 1);
        a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
        a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
        a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
        a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
        a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
        // rho + pi: b[y][2x+3y] = rol(a[x][y], r[x][y])  (lane index = x + 5 y)
        uint64_t b00 = a00, b10 = keccak_rol(a01, 1), b20 = keccak_rol(a02, 62), b05 = keccak_rol(a03, 28),
                 b15 = keccak_rol(a04, 27);
        uint64_t b16 = keccak_rol(a05, 36), b01 = keccak_rol(a06, 44), b11 = keccak_rol(a07, 6),
                 b21 = keccak_rol(a08, 55), b06 = keccak_rol(a09, 20);
        uint64_t b07 = keccak_rol(a10, 3), b17 = keccak_rol(a11, 10), b02 = keccak_rol(a12, 43),
                 b12 = keccak_rol(a13, 25), b22 = keccak_rol(a14, 39);
        uint64_t b23 = keccak_rol(a15, 41), b08 = keccak_rol(a16, 45), b18 = keccak_rol(a17, 15),
                 b03 = keccak_rol(a18, 21), b13 = keccak_rol(a19, 8);
        uint64_t b14 = keccak_rol(a20, 18), b24 = keccak_rol(a21, 2), b09 = keccak_rol(a22, 61),
                 b19 = keccak_rol(a23, 56), b04 = keccak_rol(a24, 14);
        // chi
        a00 = b00 ^ (~b01 & b02); a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00);
        a04 = b04 ^ (~b00 & b01);
        a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05);
        a09 = b09 ^ (~b05 & b06);
        a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10);
        a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15);
        a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20);
        a24 = b24 ^ (~b20 & b21);
        a00 ^= KECCAK_RC[r];
    }
    st[0] = a00; st[1] = a01; st[2] = a02; st[3] = a03; st[4] = a04; st[5] = a05; st[6] = a06; st[7] = a07;
    st[8] = a08; st[9] = a09; st[10] = a10; st[11] = a11; st[12] = a12; st[13] = a13; st[14] = a14; st[15] = a15;
    st[16] = a16; st[17] = a17; st[18] = a18; st[19] = a19; st[20] = a20; st[21] = a21; st[22] = a22; st[23] = a23;
    st[24] = a24;
}

// STROBE-128 duplex state; byte j of the sponge is byte (j & 7) of lane j >> 3 (little endian).
struct merlin_tr {
    uint64_t st[25];
    uint32_t pos, pos_begin, cur_flags;
    static const uint32_t GROUP = 1;   // threads per transcript (merlin_warp: 32)
    static const uint32_t lane = 0;    // every thread is its transcript's leader

    static const uint32_t R = 166;
    static const uint32_t F_I = 1, F_A = 2, F_C = 4, F_M = 16, F_K = 32;

    __device__ __forceinline__ void xor_byte(uint32_t j, uint32_t v) { st[j >> 3] ^= (uint64_t)(v & 0xffu) << (8 * (j & 7)); }
    __device__ __forceinline__ uint32_t take_byte(uint32_t j) {  // read and zero (PRF squeeze)
        uint32_t sh = 8 * (j & 7);
        uint32_t v = (uint32_t)(st[j >> 3] >> sh) & 0xffu;
        st[j >> 3] &= ~((uint64_t)0xff << sh);
        return v;
    }
    __device__ void run_f() {
        BPP_ASSERT(pos <= R && pos_begin <= R + 1);
        xor_byte(pos, pos_begin);
        xor_byte(pos + 1, 0x04);
        xor_byte(R + 1, 0x80);
        keccak_f1600_dev(st);
        pos = 0;
        pos_begin = 0;
    }
    // Whole Keccak lanes at a time where the sponge position allows it: one read-modify-write of the state (which lives in
    // local memory: it is indexed by the position) per 8 message bytes instead of per byte.  A 4-byte aligned message
    // is read as two 32-bit words per lane.
    __device__ void absorb(const uint8_t *d, uint32_t n) {
        while (n) {
            const uint32_t take = min(n, R - pos);
            uint32_t i = 0;
            while (i < take && ((pos + i) & 7u)) { xor_byte(pos + i, d[i]); i++; }
            if (((uintptr_t)(d + i) & 3u) == 0) {
                for (; i + 8 <= take; i += 8) {
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(d + i);
                    st[(pos + i) >> 3] ^= (uint64_t)w[0] | ((uint64_t)w[1] << 32);
                }
            } else {
                for (; i + 8 <= take; i += 8) {
                    uint64_t v = 0;
#pragma unroll
                    for (int b = 0; b < 8; b++) v |= (uint64_t)d[i + b] << (8 * b);
                    st[(pos + i) >> 3] ^= v;
                }
            }
            for (; i < take; i++) xor_byte(pos + i, d[i]);
            pos += take;
            d += take;
            n -= take;
            if (pos == R) run_f();
        }
    }
    __device__ void begin_op(uint32_t flags, bool more) {
        if (more) return;
        uint32_t old_begin = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t hdr[2] = {(uint8_t)old_begin, (uint8_t)flags};
        absorb(hdr, 2);
        if ((flags & (F_C | F_K)) && pos != 0) run_f();
    }
    __device__ void meta_ad(const uint8_t *d, uint32_t n, bool more) { begin_op(F_M | F_A, more); absorb(d, n); }
    __device__ void ad(const uint8_t *d, uint32_t n, bool more) { begin_op(F_A, more); absorb(d, n); }
    __device__ void prf(uint8_t *out, uint32_t n) {
        begin_op(F_I | F_A | F_C, false);
        for (uint32_t i = 0; i < n; i++) {
            out[i] = (uint8_t)take_byte(pos);
            if (++pos == R) run_f();
        }
    }
    // ---- merlin::Transcript ----
    __device__ void append_message(const char *label, uint32_t label_len, const uint8_t *msg, uint32_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        meta_ad((const uint8_t *)label, label_len, false);
        meta_ad(len, 4, true);
        ad(msg, n, false);
    }
    __device__ void append_u64(const char *label, uint32_t label_len, uint64_t x) {
        uint8_t b[8];
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
        append_message(label, label_len, b, 8);
    }
    __device__ void challenge_bytes(const char *label, uint32_t label_len, uint8_t *out, uint32_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        meta_ad((const uint8_t *)label, label_len, false);
        meta_ad(len, 4, true);
        prf(out, n);
    }
    // ---- TranscriptProtocol ----
    // challenge_scalar (transcript_protocol.rs:62-67): 64 bytes, Scalar::from_bytes_mod_order_wide
    // The PRF operation always starts on a fresh block (its C flag forces the permutation unless pos == 0 already), so
    // the 64 bytes are lanes 0..7 of the sponge, read and zeroed whole.
    __device__ void challenge_scalar(const char *label, uint32_t label_len, sc &out) {
        const uint8_t len[4] = {64, 0, 0, 0};
        meta_ad((const uint8_t *)label, label_len, false);
        meta_ad(len, 4, true);
        begin_op(F_I | F_A | F_C, false);   // pos == 0 afterwards
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            w[2 * i] = (uint32_t)st[i];
            w[2 * i + 1] = (uint32_t)(st[i] >> 32);
            st[i] = 0;
        }
        pos = 64;
        sc_from_wide(out, w);
    }
    // ---- state in global memory ----
    __device__ void load(const uint64_t *g) {
        for (int i = 0; i < 25; i++) st[i] = g[i];
        uint32_t m = (uint32_t)g[25];
        pos = m & 0xff; pos_begin = (m >> 8) & 0xff; cur_flags = (m >> 16) & 0xff;
    }
    __device__ void store(uint64_t *g) const {
        for (int i = 0; i < 25; i++) g[i] = st[i];
        g[25] = (uint64_t)(pos | (pos_begin << 8) | (cur_flags << 16));
    }
};

#define MERLIN_LABEL(s) (s), (uint32_t)(sizeof(s) - 1)
