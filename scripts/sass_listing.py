#!/usr/bin/env python
"""Blackwell-instruction evidence per kernel of the product library (no GPU needed):
    python scripts/sass_listing.py > profiles/r02_sass_listing.txt
Counts, per kernel of ffwm_b200/libffwm_b200.so, the SASS mnemonics of tcgen05.mma (UTC*MMA), tcgen05.ld/st (LDTM/STTM),
bulk async copies (UBLKCP), TMA tensor copies (UTMALDG/UTMASTG), mbarrier traffic (SYNCS), cp.async (LDGSTS), vector REDs."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ffwm_b200", "libffwm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
PAT = [("UTC*MMA", re.compile(r"\bUTC[A-Z]*MMA\b")), ("LDTM", re.compile(r"\bLDTM\b")), ("STTM", re.compile(r"\bSTTM\b")),
       ("UBLKCP", re.compile(r"\bUBLKCP\b")), ("UTMALDG", re.compile(r"\bUTMALDG\b")), ("UTMASTG", re.compile(r"\bUTMASTG\b")),
       ("SYNCS", re.compile(r"\bSYNCS\b")), ("LDGSTS", re.compile(r"\bLDGSTS\b")), ("RED", re.compile(r"\bRED(G)?\b")), ("HMMA", re.compile(r"\bHMMA\b"))]
counts, cur, arch = collections.OrderedDict(), None, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(.*", "", name)[:90]
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    if cur:
        for k, p in PAT:
            if p.search(line):
                counts[cur][k] += 1
print("# cuobjdump -sass ffwm_b200/libffwm_b200.so (%s): instruction counts per kernel" % arch)
print("# %-88s %s" % ("kernel", " ".join("%8s" % k for k, _ in PAT)))
for k, c in counts.items():
    if sum(c.values()):
        print("%-90s %s" % (k, " ".join("%8d" % c[n] for n, _ in PAT)))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("%-90s %s" % ("TOTAL", " ".join("%8d" % tot[n] for n, _ in PAT)))
