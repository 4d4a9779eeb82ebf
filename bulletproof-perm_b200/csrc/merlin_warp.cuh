// merlin_warp.cuh - Merlin 3.0.0 transcripts (STROBE-128 over Keccak-f[1600]) with one WARP per transcript.
//
// Same construction, labels and framing as merlin_dev.cuh (one thread per transcript; kept for the scripted test hook)
// and host_merlin.hpp, i.e. merlin::Transcript + the reference's TranscriptProtocol
// (/root/reference/bp-perm/src/transcript_protocol.rs:12-68).  Why a warp: a single transcript is a serial sponge, and with
// one thread per proof a batch of B proofs keeps B/32 warps busy - less than one per SM up to B = 4736 - so every
// Fiat-Shamir step costs the full single-thread latency of its permutations (~6 us each, profiles/r2a_launches_*.csv:
// k_tr_verify 0.40-0.54 ms, k_tr_prove_u 30 us in every inner-product round).  Here lane L < 25 holds Keccak lane L
// (x = L mod 5, y = L div 5): theta is four shuffles for the column parities and two for D, rho a per-lane rotation,
// pi one shuffle, chi two - ~45 warp instructions per round instead of ~600 thread instructions, and B warps in flight.
// All control state (pos, pos_begin) is warp-uniform; every method must be called by all 32 lanes.
#pragma once
#include <stdint.h>

#include "merlin_dev.cuh"

__device__ __constant__ const uint8_t KECCAK_ROT_LANE[32] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43, 25, 39, 41,
                                                             45, 15, 21, 8,  18, 2,  61, 56, 14, 0,  0,  0,  0,  0,  0,  0};

// message sources: byte i of the message
struct mw_mem {
    const uint8_t *p;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return p[i]; }
};
struct mw_u64 {
    uint64_t v;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return (uint32_t)(v >> (8 * i)) & 0xffu; }
};

struct merlin_warp {
    uint64_t s;                 // this lane's Keccak lane (lanes 25..31 carry zeros)
    uint32_t pos, pos_begin;    // warp-uniform
    uint32_t lane;
    uint32_t rot, pi_src, chi1, chi2, th4, th1, m5, m10, m15, m20;   // per-lane constants of the permutation

    static const uint32_t GROUP = 32;  // threads per transcript
    static const uint32_t R = 166;
    static const uint32_t F_I = 1, F_A = 2, F_C = 4, F_M = 16, F_K = 32;
    static const unsigned FULL = 0xffffffffu;

    __device__ __forceinline__ void init_lane() {
        lane = threadIdx.x & 31;
        const uint32_t l = lane < 25 ? lane : 0, x = l % 5, y = l / 5;
        rot = KECCAK_ROT_LANE[lane];
        pi_src = lane < 25 ? ((3 * y + x) % 5) + 5 * x : lane;     // B[x', y'] = rol(A[(x' + 3 y') mod 5, x'])
        chi1 = lane < 25 ? (x + 1) % 5 + 5 * y : lane;
        chi2 = lane < 25 ? (x + 2) % 5 + 5 * y : lane;
        th4 = lane < 25 ? (x + 4) % 5 : lane;
        th1 = lane < 25 ? (x + 1) % 5 : lane;
        m5 = lane < 25 ? (l + 5) % 25 : lane;
        m10 = lane < 25 ? (l + 10) % 25 : lane;
        m15 = lane < 25 ? (l + 15) % 25 : lane;
        m20 = lane < 25 ? (l + 20) % 25 : lane;
    }
    __device__ __forceinline__ static uint64_t rol(uint64_t v, uint32_t n) { return (v << n) | (v >> ((64 - n) & 63)); }

    __device__ __noinline__ void permute() {
        uint64_t a = s;
#pragma unroll 1
        for (int r = 0; r < 24; r++) {
            uint64_t c = a ^ __shfl_sync(FULL, a, m5) ^ __shfl_sync(FULL, a, m10) ^ __shfl_sync(FULL, a, m15) ^
                         __shfl_sync(FULL, a, m20);                               // column parity, in every lane of the column
            a ^= __shfl_sync(FULL, c, th4) ^ rol(__shfl_sync(FULL, c, th1), 1);   // theta
            uint64_t b = __shfl_sync(FULL, rol(a, rot), pi_src);                  // rho + pi
            a = b ^ (~__shfl_sync(FULL, b, chi1) & __shfl_sync(FULL, b, chi2));   // chi
            if (lane == 0) a ^= KECCAK_RC[r];                                     // iota
        }
        s = lane < 25 ? a : 0;
    }
    __device__ __forceinline__ void xor_byte(uint32_t j, uint32_t v) {
        if (lane == (j >> 3)) s ^= (uint64_t)(v & 0xffu) << (8 * (j & 7));
    }
    __device__ void run_f() {
        BPP_ASSERT(pos <= R && pos_begin <= R + 1);
        xor_byte(pos, pos_begin);
        xor_byte(pos + 1, 0x04);
        xor_byte(R + 1, 0x80);
        permute();
        pos = 0;
        pos_begin = 0;
    }
    // absorb n message bytes get(0..n): lane L takes the bytes that fall on sponge bytes [8L, 8L + 8)
    template <typename F>
    __device__ __forceinline__ void absorb(F get, uint32_t n) {
        uint32_t i = 0;
        while (i < n) {
            const uint32_t take = min(n - i, R - pos);
            const uint32_t lo = max(8 * lane, pos), hi = min(8 * lane + 8, pos + take);
            uint64_t w = 0;
            for (uint32_t j = lo; j < hi; j++) w |= (uint64_t)(get(i + j - pos) & 0xffu) << (8 * (j - 8 * lane));
            s ^= w;
            pos += take;
            i += take;
            if (pos == R) run_f();
        }
    }
    __device__ __forceinline__ void begin_op(uint32_t flags, bool more) {
        if (more) return;
        const uint32_t old_begin = pos_begin;
        pos_begin = pos + 1;
        mw_u64 hdr{(uint64_t)old_begin | ((uint64_t)flags << 8)};
        absorb(hdr, 2);
        if ((flags & (F_C | F_K)) && pos != 0) run_f();
    }
    template <typename F>
    __device__ __forceinline__ void meta_ad(F get, uint32_t n, bool more) { begin_op(F_M | F_A, more); absorb(get, n); }
    template <typename F>
    __device__ __forceinline__ void ad(F get, uint32_t n, bool more) { begin_op(F_A, more); absorb(get, n); }

    // ---- merlin::Transcript ----
    template <typename F>
    __device__ __forceinline__ void append(const char *label, uint32_t label_len, F get, uint32_t n) {
        meta_ad(mw_mem{(const uint8_t *)label}, label_len, false);
        meta_ad(mw_u64{(uint64_t)n}, 4, true);
        ad(get, n, false);
    }
    __device__ void append_message(const char *label, uint32_t label_len, const uint8_t *msg, uint32_t n) {
        append(label, label_len, mw_mem{msg}, n);
    }
    __device__ void append_u64(const char *label, uint32_t label_len, uint64_t x) { append(label, label_len, mw_u64{x}, 8); }
    // challenge_bytes into memory (any length); lane L writes the bytes it holds
    __device__ void challenge_bytes(const char *label, uint32_t label_len, uint8_t *out, uint32_t n) {
        meta_ad(mw_mem{(const uint8_t *)label}, label_len, false);
        meta_ad(mw_u64{(uint64_t)n}, 4, true);
        begin_op(F_I | F_A | F_C, false);
        uint32_t i = 0;
        while (i < n) {
            const uint32_t take = min(n - i, R - pos);
            const uint32_t lo = max(8 * lane, pos), hi = min(8 * lane + 8, pos + take);
            for (uint32_t j = lo; j < hi; j++) {
                const uint32_t sh = 8 * (j - 8 * lane);
                out[i + j - pos] = (uint8_t)(s >> sh);
                s &= ~((uint64_t)0xff << sh);
            }
            pos += take;
            i += take;
            if (pos == R) run_f();
        }
        __syncwarp();
    }
    // ---- TranscriptProtocol::challenge_scalar (transcript_protocol.rs:62-67): 64 bytes, wide reduction; the result is
    // computed in every lane.  The PRF operation always starts on a fresh block (its C flag forces the permutation), so
    // the 64 bytes are lanes 0..7 of the sponge.
    __device__ void challenge_scalar(const char *label, uint32_t label_len, sc &out) {
        meta_ad(mw_mem{(const uint8_t *)label}, label_len, false);
        meta_ad(mw_u64{64}, 4, true);
        begin_op(F_I | F_A | F_C, false);   // pos == 0 afterwards
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint64_t v = __shfl_sync(FULL, s, k);
            w[2 * k] = (uint32_t)v;
            w[2 * k + 1] = (uint32_t)(v >> 32);
        }
        if (lane < 8) s = 0;
        pos = 64;
        sc_from_wide(out, w);
    }
    // ---- state in global memory (same 26-word format as merlin_tr) ----
    __device__ void load(const uint64_t *g) {
        init_lane();
        s = lane < 25 ? g[lane] : 0;
        const uint32_t m = (uint32_t)g[25];
        pos = m & 0xff;
        pos_begin = (m >> 8) & 0xff;
    }
    __device__ void store(uint64_t *g) const {
        if (lane < 25) g[lane] = s;
        if (lane == 25) g[25] = (uint64_t)(pos | (pos_begin << 8));
    }
};
