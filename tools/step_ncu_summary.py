"""Summaries of an `ncu --set full` capture of ONE whole forward (267 launches at XL) made on the GPU box:

    ncu -i step.ncu-rep --page raw --csv > step_raw.csv         (on the box; the .ncu-rep itself is too large to bring back)
    python tools/step_ncu_summary.py step_raw.csv r02           -> profiles/r02_step_ncu_kernels.csv   per kernel flavour
                                                                   profiles/r02_step_ncu_kernels.md    the same, readable
                                                                   profiles/r02_gemm_traffic.json      DRAM bytes per GEMM launch (bench.py's roofline.traffic)
"""
import collections
import csv
import io
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(io.StringIO(open(raw).read())))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
M = dict(t="gpu__time_duration.sum", rd="dram__bytes_read.sum", wr="dram__bytes_write.sum",
         tp="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", dr="gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         wa="sm__warps_active.avg.pct_of_peak_sustained_active", rg="launch__registers_per_thread", gs="launch__grid_size",
         bs="launch__block_size", ia="smsp__issue_active.avg.pct")


def val(r, key):
    name = M[key]
    if name not in col:
        return float("nan")
    v = float(r[col[name]].replace(",", "") or "nan")
    u = units[col[name]]
    scale = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
             "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)
    return v * scale


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("sdp::", "")


groups = collections.OrderedDict()
for r in data:
    k = (short(r[col["Kernel Name"]]), int(val(r, "gs")), int(val(r, "bs")))
    groups.setdefault(k, []).append(r)
out = []
for (name, grid, block), rs in groups.items():
    n = len(rs)
    avg = lambda key: sum(val(r, key) for r in rs) / n
    out.append(dict(kernel=name, grid=grid, block=block, launches=n, ms=avg("t"), total_ms=avg("t") * n, dram_read_gb=avg("rd"),
                    dram_write_gb=avg("wr"), dram_pct=avg("dr"), tensor_pct=avg("tp"), warps_active_pct=avg("wa"),
                    issue_active_pct=avg("ia"), regs=int(avg("rg"))))
out.sort(key=lambda d: -d["total_ms"])
tot = sum(d["total_ms"] for d in out)
pdir = os.path.join(ROOT, "profiles")
with open(os.path.join(pdir, f"{tag}_step_ncu_kernels.csv"), "w") as f:
    keys = list(out[0].keys())
    f.write(",".join(keys) + "\n")
    for d in out:
        f.write(",".join(f'"{d[k]}"' if k == "kernel" else (f"{d[k]:.4f}" if isinstance(d[k], float) else str(d[k])) for k in keys) + "\n")
with open(os.path.join(pdir, f"{tag}_step_ncu_kernels.md"), "w") as f:
    f.write(f"# ncu --set full over one whole XL forward (batch 1024, {sum(d['launches'] for d in out)} launches), per kernel flavour ({tag})\n\n"
            "Per-launch averages; times are under ncu (cold caches, serialised, replayed): read the SHARES and the utilisations.\n"
            "DRAM % = gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed, tensor % = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active.\n\n"
            "| kernel | grid x block | launches | ms / launch | share of step | DRAM read GB | DRAM write GB | DRAM % | tensor % | issue % | regs |\n"
            "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for d in out:
        f.write(f"| `{d['kernel']}` | {d['grid']} x {d['block']} | {d['launches']} | {d['ms']:.3f} | {100 * d['total_ms'] / tot:.1f} % | "
                f"{d['dram_read_gb']:.3f} | {d['dram_write_gb']:.3f} | {d['dram_pct']:.0f} | {d['tensor_pct']:.0f} | {d['issue_active_pct']:.0f} | {d['regs']} |\n")
g = [d for d in out if "gemm_bf16_tc" in d["kernel"]]
ng = sum(d["launches"] for d in g)
traffic = {"source": f"profiles/{tag}_step_ncu_kernels.csv (ncu --set full over one XL forward, batch 1024; tools/step_ncu_summary.py)",
           "config": "XL", "batch": 1024, "gemm_launches": ng,
           "avg_dram_gb_per_launch": sum((d["dram_read_gb"] + d["dram_write_gb"]) * d["launches"] for d in g) / ng,
           "by_kernel": {f"{d['kernel']} [{d['grid']}x{d['block']}]": {"launches": d["launches"], "dram_read_gb": round(d["dram_read_gb"], 4),
                                                                     "dram_write_gb": round(d["dram_write_gb"], 4), "ms": round(d["ms"], 4)} for d in g}}
with open(os.path.join(pdir, f"{tag}_gemm_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print(f"{len(out)} flavours, {sum(d['launches'] for d in out)} launches, GEMM avg DRAM {traffic['avg_dram_gb_per_launch']:.3f} GB / launch")
