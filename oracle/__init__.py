"""oracle -- TEST INFRASTRUCTURE ONLY (see the header of oracle.c).

CPU restatement of the reference's algorithm for the dense-LA hot path, as a small C library (reference-order f32 loops,
-ffp-contract=off) with numpy glue.  May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, and by nothing under cubecl_b200/.

`oracle/_ref` (the reference compiled from its own sources) does not exist for this project: the reference is Rust and
its CPU runtime needs cargo + an LLVM bundle from the network -- unbuildable here.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "oracle.c"
LIB = HERE / "liboracle.so"

_lib = None
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    """gcc recipe for the oracle (committed here so the checker is reproducible)."""
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        cmd = ["gcc", "-O2", "-march=native", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
               str(SRC), "-lm", "-o", str(LIB)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not LIB.exists():
                raise
        L = C.CDLL(str(LIB))
        L.oracle_sum_serial_f32.restype = C.c_float
        L.oracle_sum_serial_f32.argtypes = [_f32p, C.c_size_t]
        L.oracle_sum_f64.restype = C.c_double
        L.oracle_sum_f64.argtypes = [_f32p, C.c_size_t]
        L.oracle_sum_abs_f64.restype = C.c_double
        L.oracle_sum_abs_f64.argtypes = [_f32p, C.c_size_t]
        L.oracle_sum_then_mul_f32.restype = None
        L.oracle_sum_then_mul_f32.argtypes = [_f32p, C.c_size_t, _f32p]
        L.oracle_plane_sum_f32.restype = None
        L.oracle_plane_sum_f32.argtypes = [_f32p, C.c_int, _f32p]
        L.oracle_reduce_axis_f32.restype = None
        L.oracle_reduce_axis_f32.argtypes = [C.c_int, _f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.oracle_reduce_axis_f64.restype = None
        L.oracle_reduce_axis_f64.argtypes = [C.c_int, _f32p, C.c_uint64, C.c_uint64, C.c_uint64, _f64p]
        L.oracle_matmul_f32.restype = None
        L.oracle_matmul_f32.argtypes = [_f32p, _f32p, _f32p] + [C.c_uint64] * 9
        L.oracle_matmul_f64.restype = None
        L.oracle_matmul_f64.argtypes = [_f32p, _f32p, _f64p, C.c_void_p] + [C.c_uint64] * 7
        L.oracle_matmul_points_f64.restype = None
        L.oracle_matmul_points_f64.argtypes = [_f32p, _f32p, _u64p, _u64p, C.c_uint64, _f64p, _f64p] + [C.c_uint64] * 5
        L.oracle_matmul_scaled.restype = None
        L.oracle_matmul_scaled.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, _f64p, _f64p] + [C.c_uint64] * 4
        L.oracle_num_threads.restype = C.c_int
        L.oracle_sum_blocked_f32.restype = C.c_float
        L.oracle_sum_blocked_f32.argtypes = [_f32p, C.c_size_t, C.c_int]
        L.oracle_matmul_blocked_f32.restype = None
        L.oracle_matmul_blocked_f32.argtypes = [_f32p, _f32p, _f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        _lib = L
    return _lib


OPS = {"sum": 0, "prod": 1, "max": 2, "min": 3, "argmax": 4, "argmin": 5, "mean": 6}


def _c32(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float32)


# ---------------------------------------------------------------------------------------------- reductions
def sum_serial_f32(x) -> np.float32:
    """examples/sum_things/src/lib.rs:11-18"""
    x = _c32(x).ravel()
    return np.float32(lib().oracle_sum_serial_f32(x, x.size))


def sum_f64(x) -> float:
    x = _c32(x).ravel()
    return float(lib().oracle_sum_f64(x, x.size))


def sum_abs_f64(x) -> float:
    x = _c32(x).ravel()
    return float(lib().oracle_sum_abs_f64(x, x.size))


def sum_then_mul(x) -> np.ndarray:
    """examples/sum_things/src/lib.rs:96-99"""
    x = _c32(x).ravel()
    out = np.empty_like(x)
    lib().oracle_sum_then_mul_f32(x, x.size, out)
    return out


def plane_sum(vals) -> np.ndarray:
    """vals [32, vec] -> [32, vec], every lane holding the butterfly total (shared/plane.rs:61-70)."""
    vals = _c32(vals)
    assert vals.shape[0] == 32 and vals.ndim == 2 and vals.shape[1] <= 8
    out = np.empty_like(vals)
    lib().oracle_plane_sum_f32(vals, vals.shape[1], out)
    return out


def _collapse(shape, axis):
    shape = [int(s) for s in shape]
    if axis is None:
        return 1, int(np.prod(shape, dtype=np.int64)) if shape else 1, 1, [1]
    axis %= len(shape)
    outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis else 1
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64)) if axis + 1 < len(shape) else 1
    return outer, shape[axis], inner, (shape[:axis] + shape[axis + 1:] or [1])


def reduce(x, axis, op: str) -> np.ndarray:
    """Reference-order f32 reduction along `axis` (None = all). u32 indices for arg ops."""
    x = _c32(x)
    outer, length, inner, oshape = _collapse(x.shape, axis)
    arg = op in ("argmax", "argmin")
    out = np.empty(outer * inner, dtype=np.uint32 if arg else np.float32)
    lib().oracle_reduce_axis_f32(OPS[op], x.ravel(), outer, length, inner, out.ctypes.data_as(C.c_void_p))
    return out.reshape(oshape)


def reduce_f64(x, axis, op: str) -> np.ndarray:
    x = _c32(x)
    outer, length, inner, oshape = _collapse(x.shape, axis)
    out = np.empty(outer * inner, dtype=np.float64)
    lib().oracle_reduce_axis_f64(OPS[op], x.ravel(), outer, length, inner, out)
    return out.reshape(oshape)


# ---------------------------------------------------------------------------------------------- matmul
def matmul_f32(lhs, rhs) -> np.ndarray:
    """lhs [.., M, K] @ rhs [.., K, N] (any numpy strides, batch broadcast), reference order:
    f32 `sum += l * r` over ascending k (cmma.rs:695-721).  Inputs must already be widened to f32."""
    lhs = np.asarray(lhs, dtype=np.float32)
    rhs = np.asarray(rhs, dtype=np.float32)
    if lhs.ndim != rhs.ndim or lhs.ndim < 2:
        raise ValueError("rank mismatch")
    bshape = np.broadcast_shapes(lhs.shape[:-2], rhs.shape[:-2])
    M, K = lhs.shape[-2:]
    K2, N = rhs.shape[-2:]
    if K != K2:
        raise ValueError("inner dims differ")
    lb = np.broadcast_to(lhs, bshape + (M, K))
    rb = np.broadcast_to(rhs, bshape + (K, N))
    out = np.empty(bshape + (M, N), dtype=np.float32)
    for idx in np.ndindex(*bshape):
        a = np.ascontiguousarray(lb[idx])
        b = np.ascontiguousarray(rb[idx])
        o = np.empty((M, N), dtype=np.float32)
        lib().oracle_matmul_f32(a.ravel(), b.ravel(), o.ravel(), M, N, K, K, 1, N, 1, N, 1)
        out[idx] = o
    return out


def matmul_f64(lhs, rhs):
    """(f64 result, sum_k |l||r|) for 2-D operands; the tolerance scale of SURVEY 8c."""
    a = _c32(lhs)
    b = _c32(rhs)
    M, K = a.shape
    _, N = b.shape
    out = np.empty((M, N), dtype=np.float64)
    oabs = np.empty((M, N), dtype=np.float64)
    lib().oracle_matmul_f64(a.ravel(), b.ravel(), out.ravel(), oabs.ctypes.data_as(C.c_void_p), M, N, K, K, 1, N, 1)
    return out, oabs


def matmul_scaled(lhs, rhs_nk, lhs_scales, rhs_scales, block=32):
    """Block-scaled matmul in the reference's own order (test_cmma_scaled's expected loop, cmma.rs:1572-1590).
    lhs [M,K], rhs_nk [N,K], scales widened to f32 [rows, K/block].  Returns (f32 reference order, f64 truth, sum|terms|)."""
    a, b, sa, sb = _c32(lhs), _c32(rhs_nk), _c32(lhs_scales), _c32(rhs_scales)
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and K % block == 0 and sa.shape == (M, K // block) and sb.shape == (N, K // block)
    o32 = np.empty((M, N), dtype=np.float32)
    o64 = np.empty((M, N), dtype=np.float64)
    oabs = np.empty((M, N), dtype=np.float64)
    lib().oracle_matmul_scaled(a.ravel(), b.ravel(), sa.ravel(), sb.ravel(), o32.ravel(), o64.ravel(), oabs.ravel(), M, N, K, block)
    return o32, o64, oabs


def matmul_points_f64(lhs, rhs, ms, ns):
    """f64 dot products for selected (m, n) of 2-D lhs [M,K] @ rhs [K,N] (full-size checks)."""
    a = _c32(lhs)
    b = _c32(rhs)
    K = a.shape[1]
    N = b.shape[1]
    ms = np.ascontiguousarray(ms, dtype=np.uint64)
    ns = np.ascontiguousarray(ns, dtype=np.uint64)
    out = np.empty(ms.size, dtype=np.float64)
    oabs = np.empty(ms.size, dtype=np.float64)
    lib().oracle_matmul_points_f64(a.ravel(), b.ravel(), ms, ns, ms.size, out, oabs, K, K, 1, N, 1)
    return out, oabs


# ---------------------------------------------------------------------------------------------- CPU baseline legs
def num_threads() -> int:
    return int(lib().oracle_num_threads())


def sum_blocked_f32(x, threads: int) -> np.float32:
    x = _c32(x).ravel()
    return np.float32(lib().oracle_sum_blocked_f32(x, x.size, int(threads)))


def matmul_blocked_f32(lhs, rhs_nk, threads: int) -> np.ndarray:
    a = _c32(lhs)
    b = _c32(rhs_nk)
    M, K = a.shape
    N = b.shape[0]
    out = np.empty((M, N), dtype=np.float32)
    lib().oracle_matmul_blocked_f32(a.ravel(), b.ravel(), out.ravel(), M, N, K, int(threads))
    return out
