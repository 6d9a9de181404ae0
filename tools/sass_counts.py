"""`cuobjdump -sass` of the shipped library: per kernel, how many tcgen05 MMA (UTC*MMA), TMA (UTMALDG / UTMASTG / UBLKCP),
TMEM load/store (LDTM / STTM), legacy tensor (HMMA), cp.async (LDGSTS) and packed-fp32 (FFMA2 / FADD2 / FMUL2) instructions
the SASS holds.  python tools/sass_counts.py > profiles/r02_sass_counts.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sdp-net_b200", "lib", "libsdpnet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
                               ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS"),
                               ("F*2 (packed fp32)", r"\bF(FMA|ADD|MUL)2\b"), ("MUFU.EX2", r"\bMUFU\.EX2")])
arch = re.search(r"arch = (sm_\w+)", sass)
kernels, cur = collections.OrderedDict(), None
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    kernels[cur]["instructions"] += 1 if re.match(r"\s*/\*[0-9a-f]{4}\*/", line) else 0
    for k, p in PAT.items():
        if re.search(p, line):
            kernels[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.split("\n")
print(f"# SASS instruction counts of sdp-net_b200/lib/libsdpnet_b200.so ({arch.group(1) if arch else '?'}, {len(kernels)} kernels)\n")
print("`cuobjdump -sass`, static counts per kernel (tools/sass_counts.py).  tcgen05.mma -> UTC*MMA, cp.async.bulk.tensor -> UTMALDG / "
      "UTMASTG, cp.async.bulk -> UBLKCP, tcgen05.ld / .st -> LDTM / STTM, mma.sync -> HMMA, cp.async -> LDGSTS.\n")
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print("| total | " + " | ".join(f"{k}: {tot[k]}" for k in PAT) + " |\n")
print("| kernel | SASS instrs | " + " | ".join(PAT) + " |")
print("|---|---:|" + "---:|" * len(PAT))
groups = collections.OrderedDict()
for (mangled, c), name in zip(kernels.items(), names):
    short = re.sub(r"^void\s+", "", name)
    short = re.sub(r"\(.*$", "", short).replace("sdp::", "")
    base = short.split("<")[0]
    if base == "gemm_bf16_tc_kernel":          # 80 template instances: one row per epilogue family
        t = [v.strip() for v in short[short.index("<") + 1:short.rindex(">")].split(",")]
        fam = "lean" if t[7] != "0" else "rsm" if t[8] != "0" else "headnorm" if t[3] != "0" else "generic"
        short = f"gemm_bf16_tc_kernel [{fam}, CG={t[5]}]"
        g = groups.setdefault(short, [0, collections.Counter()])
        g[0] += 1
        g[1].update(c)
        continue
    groups[short] = [1, c]
for short, (n, c) in groups.items():
    if not any(c[k] for k in list(PAT)[:8]):       # kernels without tensor-core / TMA / TMEM / cp.async instructions are left out
        continue
    label = short + (f" (x{n} instances, summed)" if n > 1 else "")
    print(f"| `{label}` | {c['instructions']} | " + " | ".join(str(c[k]) for k in PAT) + " |")
