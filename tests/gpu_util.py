"""Helpers shared by the -m gpu parity tests (all device work goes through the C ABI via cubecl_b200)."""
from __future__ import annotations

import numpy as np

import oracle
from cubecl_b200 import TensorHandle, matmul, synth

# north-star tolerances (relative to sum_k |a||b|): f32 results 1e-3, bf16/f16 results 1e-2
TOL = {"f32": 1e-3, "bf16": 1e-2, "f16": 1e-2}   # keyed by OUTPUT dtype (fp8 inputs are exact in f32 once widened)


def make_operand(shape, dtype, seed, lo=-1.0, hi=1.0, integer_mod=None):
    """(device-representation array, f32 values actually represented)"""
    n = int(np.prod(shape))
    if integer_mod:
        vals = (np.arange(n) % integer_mod).astype(np.float32).reshape(shape)
    else:
        vals = synth.uniform_f32(seed, n, lo, hi).reshape(shape)
    dev = synth.to_device_dtype(vals, dtype)
    return dev, synth.from_device_dtype(dev, dtype).reshape(shape)


def run_matmul(client, lhs_dev, rhs_dev, in_dtype, out_dtype, rhs_transposed=False, out_shape=None, lhs_transposed=False):
    """lhs_dev [..,M,K] (or [..,K,M] when lhs_transposed); rhs_dev is [..,K,N], or [..,N,K] when rhs_transposed.
    Transposed operands are passed as stride-swapped views of the same buffer (MatrixBatchLayout::MildlyPermuted)."""
    lhs = TensorHandle.from_numpy(client, lhs_dev, in_dtype)
    rhs = TensorHandle.from_numpy(client, rhs_dev, in_dtype)
    if lhs_transposed:
        lhs = lhs.transposed()
    if rhs_transposed:
        rhs = rhs.transposed()
    shape = out_shape or matmul.calculate_matmul_output(lhs.shape, rhs.shape)
    out = TensorHandle.empty_contiguous(client, shape, out_dtype)
    matmul.launch(client, lhs, rhs, out)
    client.sync()
    return synth.from_device_dtype(out.to_numpy(client), out_dtype).reshape(shape)


def check_against_oracle(got, lhs_vals, rhs_vals, out_dtype, tight=None):
    """got vs f64 ground truth, scaled by sum|a||b|; returns the max scaled error."""
    a = lhs_vals.reshape(-1, *lhs_vals.shape[-2:])
    b = rhs_vals.reshape(-1, *rhs_vals.shape[-2:])
    g = got.reshape(-1, *got.shape[-2:])
    nb = g.shape[0]
    worst = 0.0
    for i in range(nb):
        f64, fabs = oracle.matmul_f64(a[i % a.shape[0]] if a.shape[0] == 1 else a[i], b[i % b.shape[0]] if b.shape[0] == 1 else b[i])
        scale = np.maximum(fabs, 1e-30)
        err = np.abs(g[i].astype(np.float64) - f64) / scale
        worst = max(worst, float(err.max()))
    tol = tight if tight is not None else TOL[out_dtype]
    assert worst <= tol, f"max |gpu - f64| / sum|a||b| = {worst:.3e} > {tol:.1e}"
    return worst
