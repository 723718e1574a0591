#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS instructions in the built library (the table in profiles/r02_sass_counts.txt):
    python scripts/sass_counts.py [path/to/libspectral_conv_b200.so]
tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, tcgen05.st = STTM, TMA tensor load / store = UTMALDG / UTMASTG, bulk copy = UBLKCP,
tcgen05.commit = UTCBAR, mbarrier ops = SYNCS, cp.async = LDGSTS."""
import collections
import os
import re
import subprocess
import sys

OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS"]


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "neuraloperator_b200", "libspectral_conv_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), stdout=subprocess.PIPE, text=True).stdout.split("\n")
    counts, cur, it = collections.OrderedDict(), None, iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", next(it))
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
    print(f"{'kernel':54s}" + "".join(f"{o:>9s}" for o in OPS))
    for k in sorted(counts):
        print(f"{k:54s}" + "".join(f"{counts[k][o]:9d}" for o in OPS))


if __name__ == "__main__":
    main()
