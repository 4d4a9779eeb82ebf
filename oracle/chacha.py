"""Seeded RNG byte stream for the oracle: ChaCha20 keystream.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws from
`rand::thread_rng()` (circuit_lib.rs:175-176,360; lib.rs:161; weights.rs:39,59)
which cannot be seeded, so the RNG boundary is defined as a byte stream:
`rand_chacha::ChaCha20Rng::from_seed(seed32)` == DJB ChaCha20, key = seed,
64-bit block counter starting at 0, 64-bit nonce 0.  `Scalar::random` consumes
the next 64 bytes -> from_bytes_mod_order_wide; `RistrettoPoint::random`
consumes 64 bytes -> from_uniform_bytes.  Pinned by SURVEY.md E.1 (cross-checked
there against `cryptography`'s ChaCha20).
"""
from __future__ import annotations

import struct

from . import ristretto255 as R

_M32 = 0xFFFFFFFF


def _rotl(v, n):
    return ((v << n) & _M32) | (v >> (32 - n))


def _qr(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & _M32; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & _M32; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & _M32; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & _M32; s[b] = _rotl(s[b] ^ s[c], 7)


def chacha20_block(key: bytes, counter: int, nonce64: int = 0) -> bytes:
    init = list(struct.unpack("<4I", b"expand 32-byte k")) + list(struct.unpack("<8I", key))
    init += [counter & _M32, (counter >> 32) & _M32, nonce64 & _M32, (nonce64 >> 32) & _M32]
    s = list(init)
    for _ in range(10):
        _qr(s, 0, 4, 8, 12); _qr(s, 1, 5, 9, 13); _qr(s, 2, 6, 10, 14); _qr(s, 3, 7, 11, 15)
        _qr(s, 0, 5, 10, 15); _qr(s, 1, 6, 11, 12); _qr(s, 2, 7, 8, 13); _qr(s, 3, 4, 9, 14)
    return struct.pack("<16I", *[(a + b) & _M32 for a, b in zip(s, init)])


class ChaChaRng:
    """Byte stream + the dalek `random` constructors built on it."""

    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.key = bytes(seed)
        self.counter = 0
        self.buf = b""

    def fill_bytes(self, n: int) -> bytes:
        while len(self.buf) < n:
            self.buf += chacha20_block(self.key, self.counter)
            self.counter += 1
        out, self.buf = self.buf[:n], self.buf[n:]
        return out

    def scalar(self) -> int:
        return R.sc_from_wide(self.fill_bytes(64))

    def point(self):
        return R.from_uniform_bytes(self.fill_bytes(64))


def seed_from(config_id: int, index: int) -> bytes:
    """seed = LE32(config_id || index)."""
    return struct.pack("<QQ", config_id, index) + bytes(16)
