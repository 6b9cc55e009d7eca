"""`matmul::launch` surface (cubek's tiled matmul is out of tree; in-tree contract: shape rule
crates/cubecl-zspace/src/shape.rs:489-517, std-lib op convention `op::launch(client, &TensorHandle...)`
crates/cubecl-std/src/tensor/identity.rs:39-83).

out[..., m, n] = sum_k lhs[..., m, k] * rhs[..., k, n], f32 accumulation, batch dims broadcast.
The kernel behind it is the hand-written tcgen05/TMA GEMM in csrc/gemm_tcgen05.cu.
"""
from __future__ import annotations

import ctypes as C

from . import _ffi
from ._ffi import B200Error
from .client import ComputeClient, DTYPES, TensorHandle


class MatmulShapeError(ValueError):
    """calculate_matmul_output's error (shape.rs:489-517): rank mismatch or incompatible dims."""


def calculate_matmul_output(shape_lhs, shape_rhs) -> list[int]:
    """Batch-broadcast matmul shape rule, restated from crates/cubecl-zspace/src/shape.rs:489-517:
    equal rank >= 2; leading dims must be equal or one of them 1; inner dims must agree."""
    shape_lhs, shape_rhs = list(shape_lhs), list(shape_rhs)
    rank = len(shape_lhs)
    if rank != len(shape_rhs):
        raise MatmulShapeError(f"rank mismatch: lhs {rank}, rhs {len(shape_rhs)}")
    if rank < 2:
        raise MatmulShapeError("matmul needs rank >= 2")
    out = []
    for l, r in zip(shape_lhs[:-2], shape_rhs[:-2]):
        if l == r or r == 1:
            out.append(l)
        elif l == 1:
            out.append(r)
        else:
            raise MatmulShapeError(f"batch dims {l} and {r} cannot broadcast")
    if shape_lhs[-1] != shape_rhs[-2]:
        raise MatmulShapeError(f"inner dims differ: lhs k={shape_lhs[-1]}, rhs k={shape_rhs[-2]}")
    return out + [shape_lhs[-2], shape_rhs[-1]]


ACTIVATIONS = {None: 0, "none": 0, "relu": 1, "gelu": 2}


def launch(client: ComputeClient, lhs: TensorHandle, rhs: TensorHandle, out: TensorHandle, stream=None,
           alpha: float = 1.0, bias: TensorHandle | None = None, activation: str | None = None) -> None:
    """Enqueue the matmul on the client's stream.  Never raises for launch problems: errors are deferred to
    client.sync()/read_one() like the reference's launch path.
    Optional fused epilogue: out = activation(alpha * (lhs @ rhs) + bias[n]) with `bias` an f32 [N] tensor."""
    if alpha != 1.0 or bias is not None or activation not in (None, "none"):
        return _launch_fused(client, lhs, rhs, out, stream, alpha, bias, activation)
    try:
        if lhs.dtype != rhs.dtype:
            raise B200Error(6, f"lhs dtype {lhs.dtype} != rhs dtype {rhs.dtype}")
        rank = len(lhs.shape)
        if len(rhs.shape) != rank or len(out.shape) != rank:
            raise B200Error(6, "matmul: lhs, rhs and out must have equal rank")
        _ffi.check(client._lib.b200_matmul(
            client._ctx, stream, DTYPES[lhs.dtype], DTYPES[out.dtype],
            C.c_uint64(lhs.handle.ptr), C.c_uint64(rhs.handle.ptr), C.c_uint64(out.handle.ptr), rank,
            _ffi.u64_array(lhs.shape), _ffi.u64_array(lhs.strides), _ffi.u64_array(rhs.shape), _ffi.u64_array(rhs.strides),
            _ffi.u64_array(out.shape), _ffi.u64_array(out.strides)))
    except B200Error as e:
        client._defer(e)


def _launch_fused(client, lhs, rhs, out, stream, alpha, bias, activation) -> None:
    try:
        if activation not in ACTIVATIONS:
            raise B200Error(6, f"unknown activation {activation!r}")
        if bias is not None and (bias.dtype != "f32" or not bias.is_contiguous() or bias.size() != out.shape[-1]):
            raise B200Error(6, "bias must be a contiguous f32 tensor with N elements")
        if lhs.dtype != rhs.dtype:
            raise B200Error(6, f"lhs dtype {lhs.dtype} != rhs dtype {rhs.dtype}")
        rank = len(lhs.shape)
        if len(rhs.shape) != rank or len(out.shape) != rank:
            raise B200Error(6, "matmul: lhs, rhs and out must have equal rank")
        ep = _ffi.Epilogue(float(alpha), ACTIVATIONS[activation], bias.handle.ptr if bias is not None else 0)
        _ffi.check(client._lib.b200_matmul_fused(
            client._ctx, stream, DTYPES[lhs.dtype], DTYPES[out.dtype],
            C.c_uint64(lhs.handle.ptr), C.c_uint64(rhs.handle.ptr), C.c_uint64(out.handle.ptr), rank,
            _ffi.u64_array(lhs.shape), _ffi.u64_array(lhs.strides), _ffi.u64_array(rhs.shape), _ffi.u64_array(rhs.strides),
            _ffi.u64_array(out.shape), _ffi.u64_array(out.strides), C.byref(ep)))
    except B200Error as e:
        client._defer(e)


def launch_alloc(client: ComputeClient, lhs: TensorHandle, rhs: TensorHandle, out_dtype: str | None = None) -> TensorHandle:
    """Convenience: allocate `out` with the reference's shape rule, then launch."""
    shape = calculate_matmul_output(lhs.shape, rhs.shape)
    out = TensorHandle.empty_contiguous(client, shape, out_dtype or lhs.dtype)
    launch(client, lhs, rhs, out)
    return out
