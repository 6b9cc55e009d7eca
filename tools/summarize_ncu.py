"""Turn ncu artefacts from gpurun_out/ into the small tracked summaries under profiles/ (run here, no GPU needed).
usage: summarize_ncu.py <round-tag>   e.g. r01"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
PROF.mkdir(exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__waves_per_multiprocessor"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def raw(rep: Path):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return []
    H, U = rows[0], rows[1]
    return [{h: (u, v) for h, u, v in zip(H, U, r)} for r in rows[2:]]


traffic = {}
lines = [f"# ncu --set full summaries, round {tag} (source reports: gpurun_out/{tag}_*.ncu-rep, not tracked)\n"]
extra = [(p.stem[len(tag) + 1:], None) for p in sorted(OUT.glob(f"{tag}_x_*.ncu-rep"))]
for name, key in [("gemm", "gemm_bf16_8192_dram_bytes"), ("reduce", "reduce_sum_2p28_dram_bytes")] + extra:
    rep = OUT / f"{tag}_{name}.ncu-rep"
    csv_export = OUT / f"{tag}_{name}.csv"     # `ncu -i rep --page raw --csv` run on the GPU box (the .ncu-rep embeds the whole cubin)
    if rep.exists() and not (csv_export.exists() and csv_export.stat().st_mtime > rep.stat().st_mtime):
        records = raw(rep)
    elif csv_export.exists():
        rows = list(csv.reader(csv_export.open()))
        records = [{h: (u, v) for h, u, v in zip(rows[0], rows[1], r)} for r in rows[2:]] if len(rows) >= 3 else []
    else:
        continue
    for d in records[-1:]:
        kname = d.get("Kernel Name", ("", "?"))[1]
        lines.append(f"\n## {name}: {kname}\n")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:75s} {d[k][1]:>16s} {d[k][0]}\n")
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            tot = float(rd[1]) * UNIT.get(rd[0], 1.0) + float(wr[1]) * UNIT.get(wr[0], 1.0)
            if key:
                traffic[key] = tot
            lines.append(f"{'dram traffic (read+write) per launch':75s} {tot:16.0f} byte\n")
# extended captures exported remotely as CSV (the .ncu-rep embeds the whole cubin and is too large to bring back)
for csvp in sorted(OUT.glob(f"{tag}_x_*.csv")):
    rows = list(csv.reader(csvp.open()))
    if len(rows) < 3:
        continue
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(H, U, r)}
        lines.append(f"\n## {csvp.stem[len(tag) + 3:]}: {d.get('Kernel Name', ('', '?'))[1]}\n")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:75s} {d[k][1]:>16s} {d[k][0]}\n")
(PROF / f"{tag}_ncu_full_summary.txt").write_text("".join(lines))
if traffic:
    (PROF / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")

launches = OUT / f"{tag}_launches.csv"
if launches.exists():
    rows = [r for r in csv.reader(launches.open()) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, gi, bi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        a = agg.setdefault(r[ki], [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 3 --warmup 3 --quick`\n",
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n",
           f"{'kernel':44s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}  grid / block\n"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k[:44]:44s} {a[0]:8d} {a[1] / 1e3:10.1f} {a[1] / a[0] / 1e3:9.1f} {a[1] / tot * 100:6.1f}%  {a[2]} / {a[3]}\n")
    (PROF / f"{tag}_launches_summary.txt").write_text("".join(out))
    (PROF / f"{tag}_launches.csv").write_text(launches.read_text())
print("profiles written:", sorted(p.name for p in PROF.iterdir()))
