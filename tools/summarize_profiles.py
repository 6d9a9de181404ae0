"""Turn gpurun_out ncu artefacts into the small tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <launches.csv> <prof.ncu-rep> <tag>"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("sdp::", "")


if os.path.exists(launches):
    rows = []
    with open(launches) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    agg, cnt = collections.defaultdict(float), collections.Counter()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        k = short(r["Kernel Name"])
        agg[k] += ns
        cnt[k] += 1
    tot = sum(agg.values())
    with open(os.path.join(out_dir, f"{tag}_launch_list_summary.md"), "w") as f:
        f.write(f"# ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` ({tag})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` over every launch of the command "
                "(warm-up, timed, e2e and instrumented passes alike; per-launch times are cold-cache and serialised, "
                "so compare SHARES).\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{k}` | {cnt[k]} | {v / 1e6:.2f} | {100 * v / tot:.1f} % |\n")
        fam = collections.defaultdict(float)
        for k, v in agg.items():
            key = ("gemm_bf16_tc" if "gemm_bf16_tc" in k else "ln_dwconv" if "dwconv" in k else "attention" if "attention" in k
                   else "layernorm_rows" if "ln_rows" in k else "torch/other" if not k.startswith(("im2col", "fill", "pool", "tokens", "registers")) else "other sdp")
            fam[key] += v
        f.write("\n| family | share |\n|---|---:|\n")
        for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
            f.write(f"| {k} | {100 * v / tot:.1f} % |\n")
    print("wrote launch list summary:", len(cnt), "kernels", f"{tot / 1e6:.1f} ms")

if os.path.exists(rep):
    # a .csv argument is the `ncu -i rep --page raw --csv` export made on the GPU box (large reports stay there)
    raw = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    ki = hdr.index("Kernel Name")
    with open(os.path.join(out_dir, f"{tag}_ncu_full_kernels.csv"), "w") as f:
        f.write("kernel," + ",".join(f"{w} [{units[i]}]" for w, i in idx) + "\n")
        for r in data:
            f.write('"' + short(r[ki]) + '",' + ",".join(r[i].replace(",", "") for _, i in idx) + "\n")
    print("wrote ncu kernel table:", len(data), "launches")
