"""Parse the AT_TRACE printf dump of attention_tc.cu (debug builds only)."""
import collections
import re
import sys

L = [l.split() for l in open(sys.argv[1]) if re.match(r'^(tma|mma|smx) ', l)]
d = collections.OrderedDict()
for who, w, tag, t in L:
    d.setdefault((who, w), []).append((int(tag), int(t)))
for k, v in d.items():
    runs, cur = [], []
    for tag, t in v:
        if cur and t < cur[-1][1]:
            runs.append(cur)
            cur = []
        cur.append((tag, t))
    runs.append(cur)
    print(k, len(runs), 'runs; last:')
    print('  ', ' '.join(f'{a}:{b}' for a, b in runs[-1]))
