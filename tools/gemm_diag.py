"""Bring-up diagnostics for one GEMM configuration (run in its own process: a device trap kills the context).
usage: gemm_diag.py <variant> <in_dtype> <out_dtype> <rhs_t:0|1> M N K [f32mode]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from cubecl_b200 import ComputeClient  # noqa: E402
from gpu_util import make_operand, run_matmul  # noqa: E402

variant, in_dt, out_dt, rhs_t = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] == "1"
M, N, K = map(int, sys.argv[5:8])
mode = sys.argv[8] if len(sys.argv) > 8 else "tf32"
c = ComputeClient.load(0)
c.set_option("gemm.variant", variant)
c.set_option("gemm.f32", mode)
a_dev, a = make_operand((M, K), in_dt, 11)
b_dev, b = make_operand((N, K) if rhs_t else (K, N), in_dt, 12)
got = run_matmul(c, a_dev, b_dev, in_dt, out_dt, rhs_transposed=rhs_t)
bb = b.T if rhs_t else b
exp = a.astype(np.float64) @ bb.astype(np.float64)
scale = np.abs(a).astype(np.float64) @ np.abs(bb).astype(np.float64)
err = np.abs(got - exp) / np.maximum(scale, 1e-30)
tol = 1e-2 if out_dt != "f32" else (1e-3 if in_dt == "f32" else 1e-5)
bad = err > tol
tag = f"{variant} {in_dt}->{out_dt} rhs_t={int(rhs_t)} {M}x{N}x{K} {mode if in_dt == 'f32' else ''}"
if not bad.any():
    print(f"OK   {tag}: max scaled err {err.max():.3e}")
    sys.exit(0)
print(f"FAIL {tag}: max scaled err {err.max():.3e}, bad {bad.sum()}/{bad.size}")
# localise: which 128-row / 64-col blocks are wrong
rb = [(i, float(bad[i:i + 128].mean())) for i in range(0, M, 128)]
cb = [(j, float(bad[:, j:j + 64].mean())) for j in range(0, N, 64)]
print("  bad fraction per 128-row block:", [(i, round(f, 2)) for i, f in rb][:16])
print("  bad fraction per 64-col block :", [(j, round(f, 2)) for j, f in cb][:16])
idx = np.argwhere(bad)[:6]
for m, n in idx:
    print(f"  [{m},{n}] got {got[m, n]:.5f} exp {exp[m, n]:.5f}")
# does the result equal a k-prefix / a permuted-k product? (descriptor-advance bugs)
for kk in (16, 32, 64, 128):
    if kk < K:
        part = a[:, :kk].astype(np.float64) @ bb[:kk].astype(np.float64)
        if np.allclose(got, part, atol=1e-2 * max(1.0, np.abs(part).max())):
            print(f"  result equals the product over k < {kk} only")
print("  got[0,:8]", np.round(got[0, :8], 3), "\n  exp[0,:8]", np.round(exp[0, :8], 3))
sys.exit(1)
