#!/usr/bin/env python
"""Per-role (executed-count class) stall summary of an .ncu-rep source page."""
import collections, csv, io, subprocess, sys
path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; rows = rows[2:]
si, ei = hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall') and 'Not Issued' not in h]
cls = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows:
    e = int(r[ei] or 0); s = int(r[si] or 0)
    c = cls[e]; c[0] += 1; c[1] += s
    for i, h in stall: c[2][h] += int(r[i] or 0)
tot = sum(c[1] for c in cls.values())
print('total samples', tot)
for e, (n, s, cnt) in sorted(cls.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 10]:
    print(f"exec={e:8d} instrs={n:5d} samples={s:5d} ({100*s/tot:4.1f}%)", dict(cnt.most_common(5)))
