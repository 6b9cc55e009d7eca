"""Bring-up diagnostics for the block-scaled (MX) GEMM: isolates operand, scale-row, scale-k and tile effects."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from cubecl_b200 import ComputeClient, TensorHandle, matmul, synth  # noqa: E402

c = ComputeClient.load(0)


def run(kind, M, N, K, sa_bits, sb_bits, variant, a_codes=None, b_codes=None, seed=1):
    rng = np.random.default_rng(seed)
    if kind == "fp4":
        ac = rng.integers(0, 16, size=(M, K), dtype=np.uint8) if a_codes is None else a_codes
        bc = rng.integers(0, 16, size=(N, K), dtype=np.uint8) if b_codes is None else b_codes
        a_dev, b_dev = synth.pack_e2m1x2(ac), synth.pack_e2m1x2(bc)
        a, b = synth.e2m1_codes_to_f32(ac), synth.e2m1_codes_to_f32(bc)
        dt = "f4e2m1x2"
    else:
        a = synth.fp8_bits_to_f32(synth.f32_to_fp8_bits(rng.uniform(-2, 2, size=(M, K)).astype(np.float32), kind), kind)
        b = synth.fp8_bits_to_f32(synth.f32_to_fp8_bits(rng.uniform(-2, 2, size=(N, K)).astype(np.float32), kind), kind)
        a_dev, b_dev = synth.f32_to_fp8_bits(a, kind), synth.f32_to_fp8_bits(b, kind)
        dt = kind
    c.set_option("gemm.variant", variant)
    lhs, rhs = TensorHandle.from_numpy(c, a_dev, dt), TensorHandle.from_numpy(c, b_dev, dt)
    ls, rs = TensorHandle.from_numpy(c, sa_bits, "ue8m0"), TensorHandle.from_numpy(c, sb_bits, "ue8m0")
    out = TensorHandle.empty_contiguous(c, [M, N], "f32")
    matmul.launch_scaled(c, lhs, rhs, ls, rs, out)
    c.sync()
    got = out.to_numpy(c).reshape(M, N)
    _, f64, fabs = oracle.matmul_scaled(a, b, synth.ue8m0_to_f32(sa_bits), synth.ue8m0_to_f32(sb_bits), 32)
    err = np.abs(got - f64) / np.maximum(fabs, 1e-30)
    return got, f64, err


def report(tag, err, M, N):
    bad = err > 1e-4
    rows, cols = np.where(bad.any(axis=1))[0], np.where(bad.any(axis=0))[0]
    print(f"{tag:60s} max_err={np.nanmax(err):.3e} bad={int(bad.sum())}/{M * N}"
          + (f" rows[{rows.min()}..{rows.max()}] n={rows.size} cols[{cols.min()}..{cols.max()}] n={cols.size}" if bad.any() else ""), flush=True)


kinds = [a for a in sys.argv[1:] if a in ("f8e4m3", "f8e5m2", "fp4")] or ["f8e4m3", "fp4"]
variants = [a for a in sys.argv[1:] if a in ("1sm_n128", "2sm_n128", "2sm_n256", "simt")] or ["1sm_n128", "2sm_n128", "2sm_n256", "simt"]
for kind in kinds:
    for variant in variants:
        for (M, N, K) in ((128, 128, 128 if kind != "fp4" else 256), (256, 256, 512), (300, 520, 1024)):
            ns = K // 32
            ones_a, ones_b = np.full((M, ns), 127, np.uint8), np.full((N, ns), 127, np.uint8)
            try:
                _, _, e = run(kind, M, N, K, ones_a, ones_b, variant)
                report(f"{kind} {variant} {M}x{N}x{K} scales=1", e, M, N)
                ra = (120 + (np.arange(M)[:, None] % 16) + 0 * np.arange(ns)[None, :]).astype(np.uint8)
                _, _, e = run(kind, M, N, K, ra, ones_b, variant)
                report(f"{kind} {variant} {M}x{N}x{K} A scale by row", e, M, N)
                ka = (120 + 0 * np.arange(M)[:, None] + (np.arange(ns)[None, :] % 16)).astype(np.uint8)
                _, _, e = run(kind, M, N, K, ka, ones_b, variant)
                report(f"{kind} {variant} {M}x{N}x{K} A scale by k-block", e, M, N)
                rb = (120 + (np.arange(N)[:, None] % 16) + 0 * np.arange(ns)[None, :]).astype(np.uint8)
                _, _, e = run(kind, M, N, K, ones_a, rb, variant)
                report(f"{kind} {variant} {M}x{N}x{K} B scale by row", e, M, N)
                kb = (120 + 0 * np.arange(N)[:, None] + (np.arange(ns)[None, :] % 16)).astype(np.uint8)
                _, _, e = run(kind, M, N, K, ones_a, kb, variant)
                report(f"{kind} {variant} {M}x{N}x{K} B scale by k-block", e, M, N)
                rng = np.random.default_rng(7)
                _, _, e = run(kind, M, N, K, rng.integers(118, 136, size=(M, ns), dtype=np.uint8), rng.integers(118, 136, size=(N, ns), dtype=np.uint8), variant)
                report(f"{kind} {variant} {M}x{N}x{K} random scales", e, M, N)
            except Exception as ex:  # noqa: BLE001
                print(f"{kind} {variant} {M}x{N}x{K}: EXCEPTION {ex}", flush=True)
                sys.exit(1)   # a trapped kernel poisons the context: one (kind, variant) per process
