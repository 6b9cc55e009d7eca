"""SASS evidence for the headline kernels (run here, no GPU): instruction-mnemonic counts that prove the Blackwell-native
paths -- UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor loads / stores), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk),
SYNCS (mbarrier), LDG.E.128 / SHFL (the streaming reduce), HMMA (legacy mma.sync: must be absent from the GEMMs).
usage: python tools/sass_evidence.py <round-tag>   -> profiles/<tag>_sass_evidence.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "cubecl_b200" / "build"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
KERNELS = {
    "gemm_c": ["gemm_bf16_bf16_2sm_m512_kn", "gemm_e4m3_bf16_2sm_m512_kn"],
    "gemm": ["gemm_bf16_bf16_2sm_n256_kn", "gemm_tf32_f32_2sm_n256_kn", "gemm_u8_i32_2sm_n256_kn"],
    "gemm_mx": ["gemm_mxf8_bf16_2sm_n256_kk", "gemm_mxf4_bf16_2sm_n256_kk", "gemm_nvf4_bf16_2sm_n256_kk"],
    "reduce": ["reduce_all_sum_f32_tma", "reduce_all_sum_f32", "reduce_all_argmax_f32", "reduce_rows_sum_f32", "reduce_cols_sum_f32", "reduce_all_sum_f32_xgpu"],
    "aux": ["wmma_probe_bf16", "memread_probe_vec4"],
}
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCMXQMMA", "UTCCP", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "SYNCS", "HMMA",
            "LDG.E.NA.128", "LDG.E.128", "LDG", "LDS.128", "SHFL", "ATOM", "RED", "MEMBAR", "ACQBULK", "PREEXIT", "BAR.SYNC"]
out = [f"# SASS mnemonic counts of the headline kernels ({tag}): cuobjdump -sass -fun <kernel> cubecl_b200/build/<image>.cubin\n",
       "# UTC*MMA = tcgen05.mma, UTMALDG/UTMASTG = cp.async.bulk.tensor, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D), SYNCS = mbarrier,\n",
       "# ACQBULK/PREEXIT = griddepcontrol (programmatic dependent launch); HMMA (legacy mma.sync) must be 0 in every gemm_* kernel\n\n"]
for image, names in KERNELS.items():
    cubin = BUILD / f"{image}.cubin"
    for k in names:
        r = subprocess.run(["cuobjdump", "-sass", "-fun", k, str(cubin)], capture_output=True, text=True)
        ops = collections.Counter()
        variants = collections.Counter()
        total = 0
        for line in r.stdout.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            total += 1
            op = m.group(1)
            for p in PATTERNS:
                if op.startswith(p):
                    ops[p] += 1
            if op.startswith(("UTC", "UTMA", "UBLKCP", "LDTM")):
                variants[op] += 1
        out.append(f"## {k}  ({image}.cubin, {total} instructions)\n")
        out.append("   " + "  ".join(f"{p}={ops[p]}" for p in PATTERNS if ops[p]) + "\n")
        if variants:
            out.append("   forms: " + ", ".join(f"{v} x{n}" for v, n in sorted(variants.items())) + "\n")
        if k.startswith("gemm_"):
            assert ops["HMMA"] == 0, k
        out.append("\n")
(ROOT / "profiles" / f"{tag}_sass_evidence.txt").write_text("".join(out))
print("".join(out))
