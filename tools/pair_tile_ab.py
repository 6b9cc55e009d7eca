"""A/B of the GEMM tile variants on bf16 8192^3 with the power state equalised (sleep before every measurement):
default 256x256 double-accumulator tile | 512x256 pair tile (gemm.variant=2sm_m512) | 256x256 single accumulator (diagnostic)
| cuBLAS as the same-box yardstick.  CUDA events; mean of 50 back-to-back launches and mean of 1000 (power-capped regime).

usage: pair_tile_ab.py            table
       pair_tile_ab.py ncu <variant> [group_m]   a few launches of one variant (ncu target)
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cubecl_b200 import ComputeClient, TensorHandle, matmul  # noqa: E402

N = 8192
FLOPS = 2.0 * N ** 3
c = ComputeClient.load(0)
a = TensorHandle.empty_contiguous(c, [N, N], "bf16")
b = TensorHandle.empty_contiguous(c, [N, N], "bf16")
o = TensorHandle.empty_contiguous(c, [N, N], "bf16")
c.fill_uniform(a.handle, "bf16", N * N, 3, -1.0, 1.0)
c.fill_uniform(b.handle, "bf16", N * N, 4, -1.0, 1.0)
c.sync()

if len(sys.argv) > 1 and sys.argv[1] == "ncu":
    c.set_option("gemm.variant", sys.argv[2])
    if len(sys.argv) > 3:
        c.set_option("gemm.group_m", sys.argv[3])
    for _ in range(3):
        matmul.launch(c, a, b, o)
    c.sync()
    print("done", sys.argv[2])
    sys.exit(0)

import torch  # noqa: E402  (yardstick only)

dev = torch.device("cuda", 0)
ta = torch.from_numpy(a.to_numpy(c).view(np.int16).reshape(N, N)).to(dev).view(torch.bfloat16)
tb = torch.from_numpy(b.to_numpy(c).view(np.int16).reshape(N, N)).to(dev).view(torch.bfloat16)
tc = torch.empty(N, N, dtype=torch.bfloat16, device=dev)


def ours(k):
    e0, e1 = c.event(), c.event()
    c.record(e0)
    for _ in range(k):
        matmul.launch(c, a, b, o)
    c.record(e1)
    ms = c.elapsed_ms(e0, e1)
    c.sync()
    c.event_destroy(e0); c.event_destroy(e1)
    return ms


def cublas(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        torch.matmul(ta, tb, out=tc)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


# the new tile last in each pass: a fault there must not cost the other rows
CONFIGS = [("cuBLAS", None, None), ("2sm_n256 gm8", "2sm_n256", 8), ("2sm_n256a1 gm8", "2sm_n256a1", 8), ("2sm_n256 gm4", "2sm_n256", 4),
           ("2sm_m512 gm8", "2sm_m512", 8), ("2sm_m512 gm6", "2sm_m512", 6), ("2sm_m512 gm4", "2sm_m512", 4)]
res = {name: [] for name, _, _ in CONFIGS}
ref_bits = None
for rep in range(2):
    for name, variant, gm in CONFIGS:
        if variant is None:
            fn = cublas
        else:
            c.set_option("gemm.variant", variant)
            c.set_option("gemm.group_m", gm)
            fn = ours
        fn(3)
        time.sleep(2.0)
        fn(3)
        m50 = fn(50) / 50
        time.sleep(2.0)
        fn(3)
        sus = fn(1000) / 1000
        res[name].append((m50, sus))
        print(f"#{rep} {name:16s} mean of 50: {m50 * 1e3:7.1f} us {FLOPS / m50 / 1e9:7.1f} TF/s | mean of 1000: {sus * 1e3:7.1f} us "
              f"{FLOPS / sus / 1e9:7.1f} TF/s", flush=True)
        if variant is not None and rep == 0:
            bits = o.to_numpy(c)
            if ref_bits is None:
                ref_bits = bits
            else:
                print(f"   output bit-identical to 2sm_n256: {bool(np.array_equal(bits, ref_bits))}", flush=True)
base = res["2sm_n256 gm8"]
cb = res["cuBLAS"]
print("\nbest of the two passes, relative to the default tile and to cuBLAS:")
for name, _, _ in CONFIGS:
    m50 = min(r[0] for r in res[name]); sus = min(r[1] for r in res[name])
    print(f"  {name:16s} mean50 {FLOPS / m50 / 1e9:7.1f} TF/s (x{min(r[0] for r in base) / m50:.3f} of default, x{min(r[0] for r in cb) / m50:.3f} of cuBLAS)"
          f" | sustained {FLOPS / sus / 1e9:7.1f} TF/s (x{min(r[1] for r in base) / sus:.3f}, x{min(r[1] for r in cb) / sus:.3f})")
c.set_option("gemm.variant", "auto")
c.set_option("gemm.group_m", 8)
