"""Tiny ncu target: a few launches of one headline kernel (usage: prof_target.py gemm|reduce|gemm_batched|gemm_f32)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cubecl_b200 import ComputeClient, TensorHandle, matmul, reduce  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
c = ComputeClient.load(0)
if what == "reduce_rows":
    t = TensorHandle.empty_contiguous(c, [8192, 8192], "f32")
    c.fill_uniform(t.handle, "f32", 8192 * 8192, 5, 0.0, 1.0)
    out = TensorHandle.empty_contiguous(c, [8192], "f32")
    for i in range(iters):
        reduce.launch(c, t, out, 1, "sum")
elif what == "reduce_cols":
    t = TensorHandle.empty_contiguous(c, [8192, 8192], "f32")
    c.fill_uniform(t.handle, "f32", 8192 * 8192, 5, 0.0, 1.0)
    out = TensorHandle.empty_contiguous(c, [8192], "f32")
    for i in range(iters):
        reduce.launch(c, t, out, 0, "sum")
elif what == "argmax":
    n = 1 << 28
    xs = [TensorHandle.empty_contiguous(c, [n], "f32") for _ in range(2)]
    for i, x in enumerate(xs):
        c.fill_uniform(x.handle, "f32", n, 5 + i, 0.0, 1.0)
    out = TensorHandle.empty_contiguous(c, [1], "u32")
    for i in range(iters):
        reduce.launch(c, xs[i % 2], out, None, "argmax")
elif what == "reduce":
    n = 1 << 28
    xs = [TensorHandle.empty_contiguous(c, [n], "f32") for _ in range(2)]
    for i, x in enumerate(xs):
        c.fill_uniform(x.handle, "f32", n, 5 + i, 0.0, 1.0)
    out = TensorHandle.empty_contiguous(c, [1], "f32")
    for i in range(iters):
        reduce.launch(c, xs[i % 2], out, None, "sum")
elif what in ("gemm_mxf8", "gemm_mxf4"):
    import numpy as np
    n, k = 8192, 8192
    dt = "f8e4m3" if what == "gemm_mxf8" else "f4e2m1x2"
    kb = k if what == "gemm_mxf8" else k // 2
    a = TensorHandle.empty_contiguous(c, [n, kb], dt)
    b = TensorHandle.empty_contiguous(c, [n, kb], dt)
    o = TensorHandle.empty_contiguous(c, [n, n], "bf16")
    c.fill_uniform(a.handle, "f8e4m3", n * kb, 3, -1.0, 1.0)
    c.fill_uniform(b.handle, "f8e4m3", n * kb, 4, -1.0, 1.0)
    pa = TensorHandle.from_numpy(c, np.full((n // 128, k // 128, 512), 127, np.uint8), "ue8m0")
    for _ in range(iters):
        matmul.launch_scaled(c, a, b, pa, pa, o, scales_packed=True)
else:
    odt = None
    if what == "gemm_batched":
        shape, dt = [8, 4096, 4096], "bf16"
    elif what == "gemm_fp8":
        shape, dt, odt = [8192, 8192], "f8e4m3", "bf16"
    elif what == "gemm_f32_3x":
        shape, dt = [4096, 4096], "f32"
        c.set_option("gemm.f32", "3xtf32")
    elif what == "gemm_f32":
        shape, dt = [4096, 4096], "f32"
        c.set_option("gemm.f32", "tf32")
    elif what == "gemm_f32_hybrid":
        shape, dt = [4096, 4096], "f32"
        c.set_option("gemm.f32", "hybrid")
    else:
        shape, dt = [8192, 8192], "bf16"
    n = 1
    for s in shape:
        n *= s
    a = TensorHandle.empty_contiguous(c, shape, dt)
    b = TensorHandle.empty_contiguous(c, shape, dt)
    o = TensorHandle.empty_contiguous(c, shape, odt or dt)
    c.fill_uniform(a.handle, dt, n, 3, -1.0, 1.0)
    c.fill_uniform(b.handle, dt, n, 4, -1.0, 1.0)
    for _ in range(iters):
        matmul.launch(c, a, b, o)
c.sync()
print("done", what)
