"""Kernel timing sweeps on one GPU (CUDA events on the launching stream): reduce variants x grid, GEMM variants."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cubecl_b200 import ComputeClient, TensorHandle, matmul, reduce  # noqa: E402


def time_ms(c, fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    c.sync()
    e0, e1 = c.event(), c.event()
    c.record(e0)
    for _ in range(iters):
        fn()
    c.record(e1)
    ms = c.elapsed_ms(e0, e1) / iters
    c.sync()
    return ms


c = ComputeClient.load(0)
what = sys.argv[1:] or ["reduce", "gemm", "axis", "probes"]

if "reduce" in what:
    n = 1 << 28
    bufs = [TensorHandle.empty_contiguous(c, [n], "f32") for _ in range(3)]
    for i, t in enumerate(bufs):
        c.fill_uniform(t.handle, "f32", n, 5 + i, 0.0, 1.0)
    out = TensorHandle.empty_contiguous(c, [1], "f32")
    aout = TensorHandle.empty_contiguous(c, [1], "u32")
    scratch = c.empty(1024)
    k = [0]

    def run():
        k[0] += 1
        reduce.launch(c, bufs[k[0] % 3], out, None, "sum")

    def run_arg():
        k[0] += 1
        reduce.launch(c, bufs[k[0] % 3], aout, None, "argmax")

    def run_probe():
        k[0] += 1
        c.probe_memread(bufs[k[0] % 3].handle, n * 4, scratch)

    print("reduce-sum f32 2^28 (1 GiB), rotating 3 buffers (no L2 reuse); min of 3 x 20 launches:")
    rows = []
    for variant, threads, bps in (("u8", 512, 4), ("u8", 512, 2), ("u8", 256, 8), ("u8", 256, 4), ("u4", 512, 4), ("u16", 512, 2), ("u16", 256, 4),
                                  ("b8", 512, 4), ("b4", 512, 4), ("w2", 512, 4), ("w4", 512, 2), ("tma", 0, 1)):
        c.set_option("reduce.variant", variant)
        if threads:
            c.set_option("reduce.threads", threads)
            c.set_option("reduce.blocks_per_sm", bps)
        ms = min(time_ms(c, run) for _ in range(3))
        rows.append((ms, variant, threads, bps))
        print(f"  {variant:4s} threads={threads:4d} blocks/SM={bps}: {ms * 1e3:8.1f} us  {n * 4 / ms / 1e6:8.1f} GB/s   ({c.last_kernel()})", flush=True)
    c.set_option("reduce.variant", "tma")
    for stages, per_sm in ((8, 1), (6, 1), (6, 2), (4, 2), (4, 1)):
        c.set_option("reduce.tma_stages", stages)
        c.set_option("reduce.tma_ctas_per_sm", per_sm)
        for pdl in ("on", "off"):
            c.set_option("reduce.pdl", pdl)
            ms = min(time_ms(c, run) for _ in range(3))
            print(f"  tma  stages={stages} CTAs/SM={per_sm} pdl={pdl:3s}: {ms * 1e3:8.1f} us  {n * 4 / ms / 1e6:8.1f} GB/s", flush=True)
    c.set_option("reduce.pdl", "on")
    c.set_option("reduce.tma_stages", 8)
    c.set_option("reduce.tma_ctas_per_sm", 1)
    c.set_option("reduce.variant", "auto")
    c.set_option("reduce.threads", 512)
    c.set_option("reduce.blocks_per_sm", 4)
    ms_p = min(time_ms(c, run_probe) for _ in range(3))
    print(f"  reference-equivalent vec4 read probe, same buffers: {ms_p * 1e3:8.1f} us  {n * 4 / ms_p / 1e6:8.1f} GB/s")
    ms_d = min(time_ms(c, run) for _ in range(3))
    print(f"  default (auto, {c.last_kernel()}, back-to-back launches overlap through PDL): {ms_d * 1e3:8.1f} us  {n * 4 / ms_d / 1e6:8.1f} GB/s  = {ms_p / ms_d:.3f} x the probe")
    c.set_option("reduce.pdl", "off")
    ms_n = min(time_ms(c, run) for _ in range(3))
    print(f"  default with reduce.pdl=off: {ms_n * 1e3:8.1f} us  {n * 4 / ms_n / 1e6:8.1f} GB/s  = {ms_p / ms_n:.3f} x the probe")
    c.set_option("reduce.variant", "u8")
    ms_u = min(time_ms(c, run) for _ in range(3))
    c.set_option("reduce.pdl", "on")
    ms_up = min(time_ms(c, run) for _ in range(3))
    print(f"  plain 128-bit loads (u8): pdl off {ms_u * 1e3:8.1f} us {n * 4 / ms_u / 1e6:8.1f} GB/s | pdl on {ms_up * 1e3:8.1f} us {n * 4 / ms_up / 1e6:8.1f} GB/s")
    c.set_option("reduce.variant", "auto")
    ms_a = min(time_ms(c, run_arg) for _ in range(3))
    print(f"  argmax f32 2^28 ({c.last_kernel()}): {ms_a * 1e3:8.1f} us  {n * 4 / ms_a / 1e6:8.1f} GB/s")
    for dt in ("bf16", "f16"):
        tb = [TensorHandle(bufs[i].handle, [n * 2], [1], dt) for i in range(3)]   # the same gigabyte read as 2^29 16-bit elements

        def run16():
            k[0] += 1
            reduce.launch(c, tb[k[0] % 3], out, None, "max")
        ms_h = min(time_ms(c, run16) for _ in range(3))
        print(f"  max {dt} 2^29 ({c.last_kernel()}): {ms_h * 1e3:8.1f} us  {n * 4 / ms_h / 1e6:8.1f} GB/s")
    del bufs

if "gemm" in what:
    for (idt, odt, n, batch, label) in (("bf16", "bf16", 8192, 1, "bf16 8192^3"), ("bf16", "bf16", 4096, 8, "bf16 8x4096^3"),
                                        ("f32", "f32", 4096, 1, "f32 4096^3"), ("f8e4m3", "bf16", 8192, 1, "fp8 8192^3")):
        shape = [batch, n, n] if batch > 1 else [n, n]
        a = TensorHandle.empty_contiguous(c, shape, idt)
        b = TensorHandle.empty_contiguous(c, shape, idt)
        o = TensorHandle.empty_contiguous(c, shape, odt)
        c.fill_uniform(a.handle, idt, int(np.prod(shape)), 3, -1.0, 1.0)
        c.fill_uniform(b.handle, idt, int(np.prod(shape)), 4, -1.0, 1.0)
        flops = 2.0 * n * n * n * batch
        for mode in (("tf32", "3xtf32", "hybrid") if idt == "f32" else ("-",)):
            if idt == "f32":
                c.set_option("gemm.f32", mode)
            for variant in (("2sm_m512", "2sm_n256", "2sm_n128", "1sm_n128") if idt == "bf16" else
                            ("2sm_m512", "2sm_n256", "1sm_n128") if idt.startswith("f8") else ("2sm_n256", "2sm_n128", "1sm_n128")):
                for rhs_t in (False, True):
                    for gm in ((8, 4, 16) if (variant in ("2sm_n256", "2sm_m512") and not rhs_t and idt == "bf16" and batch == 1) else (8,)):
                        c.set_option("gemm.variant", variant)
                        c.set_option("gemm.group_m", gm)
                        bb = b.transposed() if rhs_t else b
                        try:
                            ms = time_ms(c, lambda: matmul.launch(c, a, bb, o), iters=10, warm=2)
                            print(f"  {label:14s} {mode:6s} {variant} rhs_t={int(rhs_t)} group_m={gm:2d}: {ms:7.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)
                        except Exception as e:  # noqa: BLE001
                            print(f"  {label} {variant} rhs_t={int(rhs_t)}: ERROR {e}")
                            raise
        c.set_option("gemm.variant", "auto")
        c.set_option("gemm.group_m", 8)
        del a, b, o

if "scaled" in what:
    # block-scaled (MX) GEMM: every tile variant, row-major scales (packing passes inside the timed call) and pre-packed
    print("block-scaled matmul, CUDA events (flops = 2 M N K):")
    for (dt, n, k) in (("f8e4m3", 8192, 8192), ("f4e2m1x2", 8192, 8192), ("f8e4m3", 4096, 4096), ("f4e2m1x2", 4096, 4096),
                       ("f4e2m1x2", 8192, 16384)):
        kb = k // 2 if dt == "f4e2m1x2" else k
        a = TensorHandle.empty_contiguous(c, [n, kb], dt)
        b = TensorHandle.empty_contiguous(c, [n, kb], dt)
        o = TensorHandle.empty_contiguous(c, [n, n], "bf16")
        c.fill_uniform(a.handle, "f8e4m3", n * kb, 3, -1.0, 1.0)   # finite e4m3 codes; as e2m1 pairs every byte is finite
        c.fill_uniform(b.handle, "f8e4m3", n * kb, 4, -1.0, 1.0)
        sa = TensorHandle.from_numpy(c, np.full((n, k // 32), 127, np.uint8), "ue8m0")
        sb = TensorHandle.from_numpy(c, np.full((n, k // 32), 126, np.uint8), "ue8m0")
        tiles, atoms = n // 128, k // 128
        pa = TensorHandle.from_numpy(c, np.full((tiles, atoms, 512), 127, np.uint8), "ue8m0")
        flops = 2.0 * n * n * k
        for variant in ("auto", "2sm_n256", "2sm_n224", "2sm_n128", "1sm_n128"):
            c.set_option("gemm.variant", variant)
            ms = min(time_ms(c, lambda: matmul.launch_scaled(c, a, b, sa, sb, o), iters=10, warm=2) for _ in range(3))
            picked = c.last_kernel()
            if variant in ("2sm_n224",):     # pre-packed scales are in the plain 128-row layout: not for the 224-wide tile
                print(f"  {dt:9s} {n}x{n}x{k} {variant}: {ms * 1e3:8.1f} us {flops / ms / 1e9:7.0f} TF/s (row-major scales, packing included)  [{picked}]", flush=True)
                continue
            msp = min(time_ms(c, lambda: matmul.launch_scaled(c, a, b, pa, pa, o, scales_packed=True), iters=10, warm=2) for _ in range(3))
            print(f"  {dt:9s} {n}x{n}x{k} {variant}: {ms * 1e3:8.1f} us {flops / ms / 1e9:7.0f} TF/s | pre-packed scales {msp * 1e3:8.1f} us {flops / msp / 1e9:7.0f} TF/s  [{picked}]", flush=True)
        c.set_option("gemm.variant", "auto")   # (rounds 1-2 left the loop's last variant, 1sm_n128, set for the nvfp4 row below)
        if dt == "f4e2m1x2":
            s16 = TensorHandle.from_numpy(c, np.full((n, k // 16), 0x38, np.uint8), "f8e4m3")
            ms = min(time_ms(c, lambda: matmul.launch_scaled(c, a, b, s16, s16, o, scale_block=16), iters=10, warm=2) for _ in range(3))
            print(f"  nvfp4     {n}x{n}x{k} auto    : {ms * 1e3:8.1f} us {flops / ms / 1e9:7.0f} TF/s (ue4m3 scale per 16, packing included)", flush=True)
        c.set_option("gemm.variant", "auto")
        del a, b, o

if "scaledab" in what:
    # block-scaled GEMM, 256 x 256 tile, pre-packed scales: scale atoms copied by the dedicated thread (default) vs by the MMA thread
    # (round-2 scheme), next to the unscaled fp8 kernel with two / one accumulator stages.  The part is power-capped and a sample
    # inherits the power state of what ran before it: 2 s pause before every 10-launch sample, configurations walked round-robin.
    import time as _time
    print("block-scaled matmul 8192^3 -> bf16, pre-packed scales; 2 s pause before every 10-launch sample, round-robin x3: min / median us, TFLOP/s of the min")
    n = k = 8192
    ops = {}
    for (tag, dt, blk) in (("mxfp8", "f8e4m3", 32), ("mxfp4", "f4e2m1x2", 32), ("nvfp4", "f4e2m1x2", 16)):
        kb = k // 2 if dt == "f4e2m1x2" else k
        a = TensorHandle.empty_contiguous(c, [n, kb], dt)
        b = TensorHandle.empty_contiguous(c, [n, kb], dt)
        c.fill_uniform(a.handle, "f8e4m3", n * kb, 3, -1.0, 1.0)
        c.fill_uniform(b.handle, "f8e4m3", n * kb, 4, -1.0, 1.0)
        pa = TensorHandle.from_numpy(c, np.full((n // 128, k // (4 * blk), 512), 127 if blk == 32 else 0x38, np.uint8), "ue8m0" if blk == 32 else "f8e4m3")
        ops[tag] = (a, b, pa, blk)
    o = TensorHandle.empty_contiguous(c, [n, n], "bf16")
    a8, b8 = ops["mxfp8"][0], ops["mxfp8"][1]
    bt = TensorHandle(b8.handle, [k, n], [1, k], "f8e4m3")
    configs = [(f"{tag} copies by {who}", ("scaled", tag, who)) for tag in ops for who in ("thread", "thread2", "mma")]
    configs += [("plain fp8 2sm_n256 (two accumulators)", ("plain", "2sm_n256", "")), ("plain fp8 2sm_n256a1 (one accumulator)", ("plain", "2sm_n256a1", ""))]
    samples = {name: [] for name, _ in configs}
    import os as _os
    for rnd in range(int(_os.environ.get("SWEEP_ROUNDS", "3"))):
        for name, (kind, x, who) in configs:
            if kind == "scaled":
                a, b, pa, blk = ops[x]
                c.set_option("gemm.variant", "auto")
                c.set_option("gemm.sf_copy", who)
                fn = lambda: matmul.launch_scaled(c, a, b, pa, pa, o, scale_block=blk, scales_packed=True)
            else:
                c.set_option("gemm.variant", x)
                fn = lambda: matmul.launch(c, a8, bt, o)
            fn(); c.sync()
            _time.sleep(2.0)
            samples[name].append(time_ms(c, fn, iters=10, warm=1))
    c.set_option("gemm.variant", "auto"); c.set_option("gemm.sf_copy", "thread")
    flops = 2.0 * n * n * k
    for name, _ in configs:
        v = sorted(samples[name])
        print(f"  {name:40s}: {v[0] * 1e3:7.1f} / {v[len(v) // 2] * 1e3:7.1f} us   {flops / v[0] / 1e9:6.0f} TFLOP/s", flush=True)
    del ops, o

if "split" in what:
    # tail split (deterministic split-K of the last partial wave): off vs forced S vs the auto policy
    print("gemm.split_k sweep (auto variant), CUDA events:")
    for (idt, odt, m, n, k, batch, mode) in (("bf16", "bf16", 4096, 4096, 4096, 1, "-"), ("bf16", "bf16", 2048, 2048, 2048, 1, "-"),
                                             ("bf16", "bf16", 1024, 1024, 8192, 1, "-"), ("bf16", "bf16", 6144, 6144, 6144, 1, "-"),
                                             ("bf16", "bf16", 5120, 5120, 5120, 1, "-"), ("bf16", "bf16", 8192, 8192, 8192, 1, "-"),
                                             ("bf16", "bf16", 512, 512, 16384, 1, "-"), ("bf16", "f32", 3072, 3072, 3072, 1, "-"),
                                             ("f32", "f32", 4096, 4096, 4096, 1, "tf32"), ("f32", "f32", 4096, 4096, 4096, 1, "3xtf32"),
                                             ("f8e4m3", "bf16", 4096, 4096, 4096, 1, "-"), ("bf16", "bf16", 4096, 4096, 4096, 3, "-")):
        sa = [batch, m, k] if batch > 1 else [m, k]
        sb = [batch, k, n] if batch > 1 else [k, n]
        so = [batch, m, n] if batch > 1 else [m, n]
        a = TensorHandle.empty_contiguous(c, sa, idt)
        b = TensorHandle.empty_contiguous(c, sb, idt)
        o = TensorHandle.empty_contiguous(c, so, odt)
        c.fill_uniform(a.handle, idt, int(np.prod(sa)), 3, -1.0, 1.0)
        c.fill_uniform(b.handle, idt, int(np.prod(sb)), 4, -1.0, 1.0)
        if idt == "f32":
            c.set_option("gemm.f32", mode)
        flops = 2.0 * m * n * k * batch
        import time as _time
        opts = ("off", "2", "3", "4", "auto")
        best = {sk: float("inf") for sk in opts}
        for rnd in range(4):                      # interleaved rounds with a pause: no option is measured in a hotter state
            for sk in (opts if rnd % 2 == 0 else opts[::-1]):
                c.set_option("gemm.split_k", sk)
                _time.sleep(0.05)
                best[sk] = min(best[sk], time_ms(c, lambda: matmul.launch(c, a, b, o), iters=10, warm=2))
        row = [f"{sk}={best[sk] * 1e3:7.1f}us/{flops / best[sk] / 1e9:6.0f}TF" for sk in opts]
        fastest = min(best[sk] for sk in opts if sk != "auto")
        row.append(f"auto/best={fastest / best['auto']:.3f}")
        print(f"  {idt:6s}->{odt:4s} {mode:6s} {batch}x{m}x{n}x{k}: " + "  ".join(row), flush=True)
        c.set_option("gemm.split_k", "auto")
        c.set_option("gemm.f32", "hybrid")
        del a, b, o

if "axis" in what:
    # the reduction tutorial's shapes (cubecl-book: 1.085 ms / 3.124 ms / 1.483 ms / 0.924 ms on an unnamed wgpu device) and
    # the shapes that stress the row / column kernels.  GB/s counts the ALGORITHMIC bytes: input read + output written.
    # Inputs <= 64 MiB are rotated over enough copies to exceed the 126 MB L2.
    print("axis reductions, CUDA events, min of 5 x 20 launches; GB/s = (input + output bytes) / time:")
    shapes = (([512, 8192], 1), ([128, 32768], 1), ([64, 256, 1024], 2), ([64, 64, 4096], 2), ([8192, 8192], 1),
              ([10000, 8192], 1), ([20000, 2048], 1), ([9000, 16384], 1), ([8192, 8192], 0), ([16, 1 << 24], 1),
              ([1 << 14, 1 << 14], 0), ([1 << 26, 4], 1), ([4, 1 << 26], 0), ([1 << 22, 64], 0), ([64, 1 << 22], 1), ([256, 1024, 1024], 1))
    sel = [a for a in what if a.startswith("vpt=")]
    vpts = [int(a.split("=")[1]) for a in sel] or [8]
    for op in ("sum", "argmax"):
        for shape, axis in shapes:
            n = int(np.prod(shape))
            copies = max(1, min(8, (192 << 20) // (n * 4) + 1))
            ts = [TensorHandle.empty_contiguous(c, shape, "f32") for _ in range(copies)]
            for i, t in enumerate(ts):
                c.fill_uniform(t.handle, "f32", n, 11 + i, 0.0, 1.0)
            oshape = reduce.output_shape(shape, axis)
            out = TensorHandle.empty_contiguous(c, oshape, reduce.output_dtype(op))
            nbytes = n * 4 + int(np.prod(oshape)) * 4
            k = [0]

            def run():
                k[0] += 1
                reduce.launch(c, ts[k[0] % copies], out, axis, op)
            res = []
            for vpt in vpts:
                c.set_option("reduce.rows_vpt", vpt)
                l0 = c.launch_count()
                run()
                launches = c.launch_count() - l0
                best = min(time_ms(c, run, iters=20, warm=3) for _ in range(5))
                res.append(f"vpt={vpt}: {best * 1e3:8.1f} us {nbytes / best / 1e6:8.1f} GB/s")
            c.set_option("reduce.rows_vpt", 8)
            print(f"  {op:6s} {str(shape):22s} axis={axis} launches={launches} x{copies} buffers: " + "  |  ".join(res), flush=True)
            del ts, out

if "f32" in what:
    print("f32 matmul on the tensor pipes (f32-equivalent 2 M N K), whole launch sequence (hybrid / 3xTF32 = 2 split passes + 1 GEMM):")
    for n in (2048, 4096, 8192):
        a = TensorHandle.empty_contiguous(c, [n, n], "f32")
        b = TensorHandle.empty_contiguous(c, [n, n], "f32")
        o = TensorHandle.empty_contiguous(c, [n, n], "f32")
        c.fill_uniform(a.handle, "f32", n * n, 1, -1.0, 1.0)
        c.fill_uniform(b.handle, "f32", n * n, 2, -1.0, 1.0)
        for mode in ("tf32", "3xtf32", "hybrid"):
            c.set_option("gemm.f32", mode)
            ms = min(time_ms(c, lambda: matmul.launch(c, a, b, o), iters=10, warm=3) for _ in range(3))
            print(f"  {n}^3 {mode:6s}: {ms * 1e3:8.1f} us  {2.0 * n ** 3 / ms / 1e9:7.1f} TFLOP/s", flush=True)
        c.set_option("gemm.f32", "hybrid")
        del a, b, o

if "launch" in what:
    # host-side cost of one launch through the Python mirror + C ABI: a reduction too small to matter on the device, issued
    # back to back; wall clock per launch (the GPU is idle most of the time)
    import time as _time
    x = TensorHandle.empty_contiguous(c, [1024], "f32")
    c.fill_uniform(x.handle, "f32", 1024, 1, 0.0, 1.0)
    out = TensorHandle.empty_contiguous(c, [1], "f32")
    a = TensorHandle.empty_contiguous(c, [128, 64], "bf16")
    b = TensorHandle.empty_contiguous(c, [64, 128], "bf16")
    o = TensorHandle.empty_contiguous(c, [128, 128], "bf16")
    for name, fn in (("reduce.launch (1024 f32)", lambda: reduce.launch(c, x, out, None, "sum")), ("matmul.launch (128x128x64 bf16)", lambda: matmul.launch(c, a, b, o))):
        for _ in range(100):
            fn()
        c.sync()
        t0 = _time.perf_counter()
        for _ in range(2000):
            fn()
        t1 = _time.perf_counter()
        c.sync()
        t2 = _time.perf_counter()
        print(f"  host cost per {name}: {(t1 - t0) / 2000 * 1e6:6.2f} us issue, {(t2 - t0) / 2000 * 1e6:6.2f} us issue + drain")

if "variants" in what:
    # validation of the static tile-variant choice: every tcgen05 variant forced against `auto`, interleaved rounds, over a
    # spread of shapes (the efficiency table the host uses was measured at 8192^3 only)
    import time as _time
    print("tile variants vs auto (bf16 -> bf16 unless noted), min over 4 interleaved rounds of 10 launches; TFLOP/s:")
    cases = ((8192, 8192, 8192, 1, "bf16", "bf16"), (4096, 4096, 4096, 1, "bf16", "bf16"), (6144, 6144, 6144, 1, "bf16", "bf16"),
             (5120, 5120, 5120, 1, "bf16", "bf16"), (3072, 3072, 3072, 1, "bf16", "bf16"), (2048, 2048, 2048, 1, "bf16", "bf16"),
             (1024, 1024, 8192, 1, "bf16", "bf16"), (512, 512, 16384, 1, "bf16", "bf16"), (8192, 1024, 4096, 1, "bf16", "bf16"),
             (1024, 8192, 4096, 1, "bf16", "bf16"), (16384, 4096, 1024, 1, "bf16", "bf16"), (4096, 4096, 4096, 8, "bf16", "bf16"),
             (2048, 2048, 2048, 16, "bf16", "bf16"), (4096, 4096, 4096, 1, "bf16", "f32"), (8192, 8192, 8192, 1, "f8e4m3", "bf16"),
             (4096, 4096, 4096, 1, "f8e4m3", "bf16"), (4096, 4096, 4096, 1, "f16", "f16"))
    worst = 1.0
    for (m, n, k, batch, idt, odt) in cases:
        sa = [batch, m, k] if batch > 1 else [m, k]
        sb = [batch, k, n] if batch > 1 else [k, n]
        so = [batch, m, n] if batch > 1 else [m, n]
        a = TensorHandle.empty_contiguous(c, sa, idt)
        b = TensorHandle.empty_contiguous(c, sb, idt)
        o = TensorHandle.empty_contiguous(c, so, odt)
        c.fill_uniform(a.handle, idt, int(np.prod(sa)), 3, -1.0, 1.0)
        c.fill_uniform(b.handle, idt, int(np.prod(sb)), 4, -1.0, 1.0)
        flops = 2.0 * m * n * k * batch
        opts = ["auto", "2sm_n256", "1sm_n128"] + (["2sm_m512"] if odt != "f32" else []) + (["2sm_n128"] if not idt.startswith("f8") else [])
        best = {v: float("inf") for v in opts}
        picked = ""
        for rnd in range(4):
            for v in (opts if rnd % 2 == 0 else opts[::-1]):
                c.set_option("gemm.variant", v)
                _time.sleep(0.03)
                best[v] = min(best[v], time_ms(c, lambda: matmul.launch(c, a, b, o), iters=10, warm=2))
                if v == "auto":
                    picked = c.last_kernel()
        c.set_option("gemm.variant", "auto")
        fastest = min(best[v] for v in opts if v != "auto")
        worst = min(worst, fastest / best["auto"])
        print(f"  {idt}->{odt} {batch}x{m}x{n}x{k}: " + "  ".join(f"{v}={flops / best[v] / 1e9:6.0f}" for v in opts) +
              f"  | auto/best={fastest / best['auto']:.3f} ({picked})", flush=True)
        del a, b, o
    print(f"  worst auto/best over the sweep: {worst:.3f}")

if "cols" in what:
    # column (outer-axis) reductions: how many blocks a segmented axis should aim at (reduce.cols_split_target x SMs)
    print("column reductions, GB/s = (input + output bytes) / time, min of 5 x 20 launches:")
    for shape, axis in (([8192, 8192], 0), ([1 << 14, 1 << 14], 0), ([1 << 22, 64], 0), ([4096, 4096], 0), ([64, 4096, 1024], 1)):
        n = int(np.prod(shape))
        copies = max(1, min(8, (192 << 20) // (n * 4) + 1))
        ts = [TensorHandle.empty_contiguous(c, shape, "f32") for _ in range(copies)]
        for i, t in enumerate(ts):
            c.fill_uniform(t.handle, "f32", n, 11 + i, 0.0, 1.0)
        for op in ("sum", "argmax"):
            oshape = reduce.output_shape(shape, axis)
            out = TensorHandle.empty_contiguous(c, oshape, reduce.output_dtype(op))
            nbytes = n * 4 + int(np.prod(oshape)) * 4
            k = [0]

            def run():
                k[0] += 1
                reduce.launch(c, ts[k[0] % copies], out, axis, op)
            res = []
            for target in (0, 4, 8, 12, 16):
                c.set_option("reduce.cols_split_target", target)
                best = min(time_ms(c, run, iters=20, warm=3) for _ in range(4))
                res.append(f"t{target}:{best * 1e3:6.1f}us/{nbytes / best / 1e6:5.0f}")
            c.set_option("reduce.cols_split_target", "")
            print(f"  {op:6s} {str(shape):20s} axis={axis}: " + " ".join(res), flush=True)
        del ts

if "unaligned" in what:
    # operands TMA cannot describe are staged (one copy pass) and run on the tensor cores: K = 4097 against K = 4096
    print("bf16 4096 x 4096 x K, lhs [M,K] row-major (row pitch 2K bytes), rhs [K,N] row-major:")
    for k_ in (4096, 4097, 4104):
        a = TensorHandle.empty_contiguous(c, [4096, k_], "bf16")
        b = TensorHandle.empty_contiguous(c, [k_, 4096], "bf16")
        o = TensorHandle.empty_contiguous(c, [4096, 4096], "bf16")
        c.fill_uniform(a.handle, "bf16", 4096 * k_, 3, -1.0, 1.0)
        c.fill_uniform(b.handle, "bf16", 4096 * k_, 4, -1.0, 1.0)
        l0 = c.launch_count()
        matmul.launch(c, a, b, o)
        launches = c.launch_count() - l0
        ms = min(time_ms(c, lambda: matmul.launch(c, a, b, o), iters=20, warm=3) for _ in range(3))
        print(f"  K={k_}: {ms * 1e3:8.1f} us  {2.0 * 4096 * 4096 * k_ / ms / 1e9:7.0f} TFLOP/s  launches={launches}  last kernel {c.last_kernel()}", flush=True)
        del a, b, o

if "probes" in what:
    scratch = c.empty(1024)
    for dt in ("f16", "bf16"):
        n_iter = 4096
        ops = [0.0]

        def run():
            ops[0] = c.probe_wmma(dt, n_iter, scratch)

        ms = time_ms(c, run, iters=5, warm=2)
        print(f"reference-equivalent wmma probe ({dt}, 16x16x16, grid SMs*32 x 256): {ops[0] / ms / 1e9:8.1f} TFLOP/s")
    ops = [0.0]

    def run_u():
        ops[0] = c.probe_umma(8192, scratch)

    ms = time_ms(c, run_u, iters=5, warm=2)
    import numpy as _np
    vals = _np.frombuffer(c.read_one(scratch), dtype=_np.float32)[:4]
    print(f"tcgen05 peak probe (UMMA 256x256x16 bf16, 74 CTA pairs, smem-resident): {ops[0] / ms / 1e9:8.1f} TFLOP/s  acc[0]={vals[0]} (expect {64 * 8192})")
    for dt, scaled, label in (("f8e4m3", False, "kind::f8f6f4 e4m3"), ("f8e4m3", True, "kind::mxf8f6f4.block_scale e4m3"),
                              ("f4e2m1x2", True, "kind::mxf4.block_scale e2m1")):
        def run_k():
            ops[0] = c.probe_umma_kind(dt, scaled, 8192, scratch)

        ms = time_ms(c, run_k, iters=5, warm=2)
        vals = _np.frombuffer(c.read_one(scratch), dtype=_np.float32)[:4]
        print(f"tcgen05 peak probe ({label}, UMMA 256x256, smem-resident): {ops[0] / ms / 1e9:8.1f} TFLOP/s  acc[0]={vals[0]}")
    buf = c.empty(512 << 20)
    c.fill_modulo(buf, "f32", (512 << 20) // 4, 8)
    ms = time_ms(c, lambda: c.probe_memread(buf, 512 << 20, scratch), iters=10, warm=2)
    print(f"reference-equivalent vec4 read probe (512 MiB): {(512 << 20) / ms / 1e6:8.1f} GB/s")
