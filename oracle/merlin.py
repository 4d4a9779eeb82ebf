"""Merlin 3.0.0 transcript (STROBE-128 over Keccak-f[1600]) oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  merlin 3.0.0 is pinned at
/root/reference/bp-perm/Cargo.lock:189-192 and is not vendored; this restates
its published construction.  Pinned by the Merlin KAT (tests/test_oracle.py).
Reference call sites: /root/reference/bp-perm/src/transcript_protocol.rs:26-68
and lib.rs:172 (`Transcript::new(b"test")`).
"""
from __future__ import annotations

import struct

from . import ristretto255 as R

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [
    [0, 36, 3, 41, 18],
    [1, 44, 10, 45, 2],
    [62, 6, 43, 15, 61],
    [28, 55, 25, 21, 56],
    [27, 20, 39, 8, 14],
]
_M = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M if n else v


def keccak_f1600(lanes):
    """lanes: list of 25 u64, index x + 5*y."""
    a = lanes
    for rc in _RC:
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], _ROT[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= rc
    return a


_STROBE_R = 166
_FLAG_I, _FLAG_A, _FLAG_C, _FLAG_T, _FLAG_M, _FLAG_K = 1, 2, 4, 8, 16, 32


class Strobe128:
    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, _STROBE_R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        self.state = st
        self._permute()
        self.pos = 0
        self.pos_begin = 0
        self.cur_flags = 0
        self.meta_ad(protocol_label, False)

    def clone(self):
        o = object.__new__(Strobe128)
        o.state = bytearray(self.state)
        o.pos, o.pos_begin, o.cur_flags = self.pos, self.pos_begin, self.cur_flags
        return o

    def _permute(self):
        lanes = list(struct.unpack("<25Q", bytes(self.state)))
        self.state = bytearray(struct.pack("<25Q", *keccak_f1600(lanes)))

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[_STROBE_R + 1] ^= 0x80
        self._permute()
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data: bytes):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == _STROBE_R:
                self._run_f()

    def _squeeze(self, n: int) -> bytes:
        out = bytearray(n)
        for i in range(n):
            out[i] = self.state[self.pos]
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == _STROBE_R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags: int, more: bool):
        if more:
            assert self.cur_flags == flags
            return
        assert not (flags & _FLAG_T)
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        force_f = bool(flags & (_FLAG_C | _FLAG_K))
        if force_f and self.pos != 0:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool):
        self._begin_op(_FLAG_M | _FLAG_A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool):
        self._begin_op(_FLAG_A, more)
        self._absorb(data)

    def prf(self, n: int, more: bool) -> bytes:
        self._begin_op(_FLAG_I | _FLAG_A | _FLAG_C, more)
        return self._squeeze(n)


class Transcript:
    """merlin::Transcript plus the reference's TranscriptProtocol extension
    (transcript_protocol.rs:26-68)."""

    def __init__(self, label: bytes):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def clone(self):
        o = object.__new__(Transcript)
        o.strobe = self.strobe.clone()
        return o

    def append_message(self, label: bytes, message: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", len(message)), True)
        self.strobe.ad(message, False)

    def append_u64(self, label: bytes, x: int):
        self.append_message(label, struct.pack("<Q", x))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", n), True)
        return self.strobe.prf(n, False)

    # ---- TranscriptProtocol (transcript_protocol.rs) ----
    def arithmetic_domain_sep(self, n: int):  # :27-30
        self.append_message(b"dom-sep", b"acp v1")
        self.append_u64(b"n", n)

    def append_scalar(self, label: bytes, s: int):  # :32-34
        self.append_message(label, R.sc_bytes(s))

    def append_vec_scalar(self, label: bytes, scalars):  # :36-43
        self.append_message(label, encode_vec_scalar(scalars))

    def append_point(self, label: bytes, compressed: bytes):  # :45-47
        assert len(compressed) == 32
        self.append_message(label, compressed)

    def validate_and_append_point(self, label: bytes, compressed: bytes) -> bool:  # :48-60
        if compressed == bytes(32):
            return False
        self.append_message(label, compressed)
        return True

    def challenge_scalar(self, label: bytes) -> int:  # :62-67
        return R.sc_from_wide(self.challenge_bytes(label, 64))


def encode_vec_scalar(scalars) -> bytes:
    """`Vec<String>.encode::<u64>()` of bytevec 0.2.0 applied to the decimal
    strings of `I256::from_le_bytes(scalar)` (transcript_protocol.rs:36-43).

    bytevec's collection encoding (un-vendored; restated from its published
    source): a u64 big-endian-less... the crate writes sizes with
    `BVSize::encode` in the host's *little-endian*-independent form: each
    element is prefixed by its byte length, and the whole list by the total
    payload length, all as 8-byte big-endian integers.  No challenge is drawn
    after the "l"/"r" appends in the reference flow (circuit_lib.rs:464-468),
    so this encoding never influences any observable output; it is kept
    host-side and identical between the oracle and the product.
    """
    parts = []
    for s in scalars:
        v = s % R.L  # canonical scalars are < 2^253, so I256 is non-negative
        d = str(v).encode()
        parts.append(struct.pack(">Q", len(d)) + d)
    body = b"".join(parts)
    return struct.pack(">Q", len(body)) + body
