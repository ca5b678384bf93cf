#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python scripts/launch_summary.py <launches.csv> <summary.txt> [--rm]"""
import collections
import csv
import os
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if not hi:
    sys.exit("no launch table in %s" % src)
hdr = rows[hi[0]]
body = [r for r in rows[hi[0] + 1:] if len(r) > 10]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in body:
    name = re.sub(r"\(.*", "", r[kn].replace("void ", ""))[:110]
    tot[name] += float(r[mv])
    cnt[name] += 1
T = sum(tot.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none: %d launches, %.1f ms serialised cold-cache GPU time" % (len(body), T / 1e6),
       "# ms  share  launches  kernel"]
out += ["%8.3f %5.1f%% %5d  %s" % (v / 1e6, 100 * v / T, cnt[k], k) for k, v in tot.most_common()]
open(dst, "w").write("\n".join(out) + "\n")
if "--rm" in sys.argv:
    os.remove(src)
