"""Compact summary of an `ncu --page source --csv` export: stall mix and the hottest SASS lines of one kernel.
python tools/ncu_source_summary.py <source.csv> <out.md> [title]"""
import io
import sys

import pandas as pd

src, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
lines = open(src).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
kernel = next((l for l in lines[:start] if "Kernel Name" in l), "")
df = pd.read_csv(io.StringIO("\n".join(lines[start:])), dtype=str)
df = df[df["Address"].str.startswith("0x", na=False)].reset_index(drop=True)
first = df["Address"].iloc[0]
idx = df.index[df["Address"] == first].tolist()
if len(idx) > 1:
    df = df.iloc[:idx[1]].copy()                      # first kernel of the export only
cols = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
for c in cols + ["# Samples", "Instructions Executed"]:
    df[c] = pd.to_numeric(df[c], errors="coerce").fillna(0)
tot = df["# Samples"].sum()
with open(out, "w") as f:
    f.write(f"# {title}\n\n{kernel[:300]}\n\n{len(df)} SASS instructions, {int(tot)} warp samples, "
            f"{int(df['Instructions Executed'].sum())} warp instructions executed.\n\n| stall reason | share of samples |\n|---|---:|\n")
    for k, v in (df[cols].sum().sort_values(ascending=False) / tot).head(10).items():
        f.write(f"| {k[6:]} | {100 * v:.1f} % |\n")
    f.write("\n| samples | executed | SASS | top stalls |\n|---:|---:|---|---|\n")
    for _, r in df.sort_values("# Samples", ascending=False).head(25).iterrows():
        st = sorted([(int(r[c]), c[6:]) for c in cols], reverse=True)[:2]
        f.write(f"| {int(r['# Samples'])} | {int(r['Instructions Executed'])} | `{r.Source.strip()[:80]}` | "
                f"{', '.join(f'{n} {c}' for c, n in [(b, a) for a, b in st])} |\n")
print("wrote", out)
