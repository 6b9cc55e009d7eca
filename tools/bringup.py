"""First-contact script for a GPU box: staged checks, each stage in its own process with a timeout, so one trap or hang
does not hide the rest.  Writes everything to stdout (redirect into gpurun_out/)."""
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PY = sys.executable


def stage(name, cmd, timeout):
    t0 = time.time()
    print(f"\n===== {name} :: {' '.join(cmd)}", flush=True)
    try:
        r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
        out, rc = r.stdout, r.returncode
    except subprocess.TimeoutExpired as e:
        out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT"
    tail = "\n".join(out.splitlines()[-60:])
    print(tail)
    print(f"===== {name}: rc={rc} ({time.time() - t0:.1f}s)", flush=True)
    return rc == 0


what = sys.argv[1:] or ["runtime", "reduce", "simt", "diag", "matmul", "perf"]
subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total", "--format=csv"], cwd=ROOT)
if "runtime" in what:
    stage("runtime", [PY, "-m", "pytest", "tests/test_runtime_gpu.py", "-x", "-q", "-m", "gpu"], 300)
if "reduce" in what:
    stage("reduce", [PY, "-m", "pytest", "tests/test_reduce_gpu.py", "-q", "-m", "gpu", "-x"], 600)
if "simt" in what:
    stage("simt", [PY, "-m", "pytest", "tests/test_matmul_gpu.py", "-q", "-m", "gpu", "-k", "simt"], 300)
if "diag" in what:
    for variant in ("1sm_n128", "2sm_n128", "2sm_n256"):
        for rhs_t in ("1", "0"):
            for (idt, odt) in (("bf16", "f32"), ("bf16", "bf16"), ("f32", "f32")):
                for (M, N, K) in ((128, 128, 64), (256, 256, 64), (384, 512, 320)):
                    stage(f"diag {variant}", [PY, "tools/gemm_diag.py", variant, idt, odt, rhs_t, str(M), str(N), str(K)], 90)
if "diagf32" in what:
    for variant in ("1sm_n128", "2sm_n256"):
        for mode in ("tf32", "3xtf32"):
            for (M, N, K) in ((128, 128, 64), (384, 512, 320)):
                stage(f"diagf32 {variant}", [PY, "tools/gemm_diag.py", variant, "f32", "f32", "0", str(M), str(N), str(K), mode], 90)
if "matmul" in what:
    stage("matmul", [PY, "-m", "pytest", "tests/test_matmul_gpu.py", "-q", "-m", "gpu"], 900)
if "perf" in what:
    stage("perf", [PY, "tools/perf_sweep.py"], 900)
if "smoke" in what:
    stage("smoke", [PY, "__graft_entry__.py", "smoke"], 300)
if "bench" in what:
    stage("bench", [PY, "bench.py"], 900)
    stage("bench-ref", [PY, "bench.py", "--impl", "reference", "--steps", "3", "--warmup", "1"], 600)
