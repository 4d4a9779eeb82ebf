"""Python mirror of the reference's util.rs / poly.rs operator layer, executed on the GPU through the
C ABI.  Same names, argument meaning and error behaviour (a Rust `panic!` on mismatched dimensions
becomes a ValueError).  Scalars are ints (or 32-byte little-endian bytes); results are ints."""
from __future__ import annotations

import ctypes

from . import _lib
from .backend import Backend, scalars_to_bytes


def _ints(buf: bytes):
    return [int.from_bytes(buf[i:i + 32], "little") for i in range(0, len(buf), 32)]


class Ops:
    def __init__(self, backend: Backend):
        self.be = backend
        self.lib = backend._lib
        self.ctx = backend._ctx

    def _chk(self, rc, what):
        if rc == -4:
            raise ValueError(f"{what}: dimension mismatch (the reference panics here)")
        self.be._check(rc)

    def inner_product(self, a, b):  # util.rs:84-94
        out = ctypes.create_string_buffer(32)
        self._chk(self.lib.bpp_inner_product(self.ctx, scalars_to_bytes(a), len(a), scalars_to_bytes(b), len(b), out),
                  "inner_product(a,b)")
        return int.from_bytes(out.raw, "little")

    def hadamard_V(self, a, b):  # util.rs:6-20
        out = ctypes.create_string_buffer(32 * max(len(a), 1))
        self._chk(self.lib.bpp_hadamard_V(self.ctx, scalars_to_bytes(a), len(a), scalars_to_bytes(b), len(b), out),
                  "hadamard_V(a, b)")
        return _ints(out.raw[:32 * len(a)])

    def vm_mult(self, a, b):  # util.rs:22-38
        rows, cols = len(b), len(b[0])
        flat = b"".join(scalars_to_bytes(r) for r in b)
        out = ctypes.create_string_buffer(32 * rows)
        self._chk(self.lib.bpp_vm_mult(self.ctx, scalars_to_bytes(a), len(a), flat, rows, cols, out), "vm_mult(a,b)")
        return _ints(out.raw)

    def mv_mult(self, a, b):  # util.rs:40-56
        rows, cols = len(a), len(a[0])
        flat = b"".join(scalars_to_bytes(r) for r in a)
        out = ctypes.create_string_buffer(32 * cols)
        self._chk(self.lib.bpp_mv_mult(self.ctx, flat, rows, cols, scalars_to_bytes(b), len(b), out), "mv_mult(a,b)")
        return _ints(out.raw)

    def lm_mult(self, a, b):  # util.rs:58-61
        return self.vm_mult(list(a), b)

    def exp_iter(self, x, count):  # util.rs:63-65,139-157 (.take(count))
        out = ctypes.create_string_buffer(32 * max(count, 1))
        self._chk(self.lib.bpp_exp_iter(self.ctx, scalars_to_bytes([x]), count, out), "exp_iter")
        return _ints(out.raw[:32 * count])

    def scalar_powers(self, x, first, count):
        out = ctypes.create_string_buffer(32 * max(count, 1))
        self._chk(self.lib.bpp_scalar_powers(self.ctx, scalars_to_bytes([x]), first, count, out), "scalar_powers")
        return _ints(out.raw[:32 * count])

    def scalar_exp(self, x, pow_):  # util.rs:67-82
        out = ctypes.create_string_buffer(32)
        self._chk(self.lib.bpp_scalar_exp(self.ctx, scalars_to_bytes([x]), pow_, out), "scalar_exp")
        return int.from_bytes(out.raw, "little")

    def invert_all(self, a):  # circuit_lib.rs:273-275
        out = ctypes.create_string_buffer(32 * max(len(a), 1))
        self._chk(self.lib.bpp_scalar_invert(self.ctx, scalars_to_bytes(a), len(a), out), "invert")
        return _ints(out.raw[:32 * len(a)])

    def from_bytes_mod_order_wide(self, blobs: bytes):  # transcript_protocol.rs:62-67, Scalar::random
        n = len(blobs) // 64
        out = ctypes.create_string_buffer(32 * max(n, 1))
        self._chk(self.lib.bpp_scalar_from_wide(self.ctx, blobs, n, out), "from_bytes_mod_order_wide")
        return _ints(out.raw[:32 * n])

    def reduce_scalars(self, raw: bytes):  # traits.rs:7-17
        n = len(raw) // 32
        out = ctypes.create_string_buffer(32 * max(n, 1))
        self._chk(self.lib.bpp_scalar_reduce(self.ctx, raw, n, out), "reduce_scalars")
        return _ints(out.raw[:32 * n])

    # poly.rs
    def special_inner_product(self, lhs, rhs):  # lhs, rhs: 4 lists of n scalars -> [t1..t6]
        n = len(lhs[0])
        out = ctypes.create_string_buffer(192)
        self._chk(self.lib.bpp_vecpoly3_special_inner_product(
            self.ctx, b"".join(scalars_to_bytes(v) for v in lhs), b"".join(scalars_to_bytes(v) for v in rhs), n, out),
            "special_inner_product")
        return _ints(out.raw)

    def vecpoly3_eval(self, coeffs, x):
        n = len(coeffs[0])
        out = ctypes.create_string_buffer(32 * n)
        self._chk(self.lib.bpp_vecpoly3_eval(self.ctx, b"".join(scalars_to_bytes(v) for v in coeffs), n,
                                             scalars_to_bytes([x]), out), "VecPoly3::eval")
        return _ints(out.raw)

    def poly6_eval(self, t, x):
        out = ctypes.create_string_buffer(32)
        self._chk(self.lib.bpp_poly6_eval(self.ctx, scalars_to_bytes(t), scalars_to_bytes([x]), out), "Poly6::eval")
        return int.from_bytes(out.raw, "little")
