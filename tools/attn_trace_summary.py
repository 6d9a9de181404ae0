"""Summarise a cycle trace written by tools/attn_trace.sh: per role/warp, the trace codes with deltas."""
import collections
import sys


def main(path, warps=("smx:w2", "smx:w6", "mma:w1", "prd:w0"), first=40, count=110):
    ev = collections.defaultdict(list)
    for line in open(path):
        p = line.split()
        if len(p) == 4 and p[2].isdigit():
            ev[f"{p[0]}:{p[1]}"].append((int(p[2]), int(p[3])))
    t0 = min(v[0][1] for v in ev.values())
    for k in warps:
        v = ev.get(k, [])
        out, prev = [], None
        for c, t in v[first:first + count]:
            out.append(f"{c}@{t - t0}" + (f"(+{t - prev})" if prev else ""))
            prev = t
        print(k, len(v))
        print(" ".join(out))


if __name__ == "__main__":
    main(sys.argv[1], *([tuple(sys.argv[2].split(","))] if len(sys.argv) > 2 else []))
