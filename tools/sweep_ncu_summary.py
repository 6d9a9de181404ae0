"""Joins the ncu launch list of `tools/layer_sweep.py --once` (csv with gpu__time_duration, dram__throughput and
sm__pipe_tensor_cycles_active per launch) with gpurun_out/sweep_manifest.json and prints, per grid / batch / layer, the
layer's time and its time-weighted DRAM and tensor-pipe utilisation, plus the same per kernel family.
python tools/sweep_ncu_summary.py <launches.csv> <manifest.json> > profiles/r02_layer_sweep_ncu.md"""
import csv
import io
import json
import sys
from collections import defaultdict

T, D, P = "gpu__time_duration.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"


def family(name):
    for key, fam in (("gemm_bf16_tc", "gemm"), ("attention", "attention"), ("ln_dwconv_slab", "dwconv"), ("token_stats", "stats"),
                     ("row_stats", "stats"), ("ln_rows", "layernorm")):
        if key in name:
            return fam
    return "other"


def main():
    text = open(sys.argv[1]).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    launches = defaultdict(dict)          # id -> {metric: value, name}
    order = []
    for r in rows:
        if "sdp::" not in r["Kernel Name"] or "row_stats_kernel" in r["Kernel Name"]:    # (the sweep's own set-up pass)
            continue
        i = int(r["ID"])
        if i not in launches:
            order.append(i)
        launches[i]["name"] = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == T:
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"].replace("second", "s").replace("nsecond", "ns"), 1e-6)
        launches[i][r["Metric Name"]] = v
    manifest = json.load(open(sys.argv[2]))
    pos = 0
    print("# Layer sweep with ncu counters (BASELINE.json configs[4]): embed 768, 8 heads, 5 registers, bf16\n")
    print("One call per (grid, batch, layer) after a warm-up call, `ncu --metrics gpu__time_duration.sum,dram__throughput...,"
          "sm__pipe_tensor_cycles_active... --clock-control none` (per-launch times are cold-cache and serialised: read the "
          "shares and the utilisations, the timing table is `r02_layer_sweep.md`).  DRAM % = dram__throughput.avg.pct_of_peak_sustained_elapsed, "
          "tensor % = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed, both time-weighted over the layer's launches; "
          "per family: share of the layer's time / DRAM % / tensor %.\n")
    tables = defaultdict(list)
    for m in manifest:
        ids = order[pos:pos + m["launches"]]
        pos += m["launches"]
        if m["pass"] != "meas":
            continue
        tot = sum(launches[i][T] for i in ids)
        fam = defaultdict(lambda: [0.0, 0.0, 0.0])
        for i in ids:
            f = fam[family(launches[i]["name"])]
            f[0] += launches[i][T]
            f[1] += launches[i][T] * launches[i].get(D, 0.0)
            f[2] += launches[i][T] * launches[i].get(P, 0.0)
        d = sum(f[1] for f in fam.values()) / tot
        p = sum(f[2] for f in fam.values()) / tot
        fs = "; ".join(f"{k} {100 * v[0] / tot:.0f}% / {v[1] / v[0]:.0f} / {v[2] / v[0]:.0f}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0]))
        tables[(m["grid"], m["layer"])].append(f"| {m['batch']} | {len(ids)} | {tot:.3f} | {d:.1f} | {p:.1f} | {fs} |")
    for (grid, layer), lines in tables.items():
        print(f"\n## {layer}, {grid} x {grid} grid\n")
        print("| batch | launches | ms (ncu, serialised) | DRAM % | tensor % | per family (share / DRAM % / tensor %) |")
        print("|---:|---:|---:|---:|---:|---|")
        print("\n".join(lines))
    assert pos == len(order), (pos, len(order))


if __name__ == "__main__":
    main()
