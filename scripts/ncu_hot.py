#!/usr/bin/env python
"""Per-address-bucket and top-instruction stall samples of one kernel in an .ncu-rep.
    python scripts/ncu_hot.py <rep> <kernel-regex> [bucket_bytes]"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3], 0) if len(sys.argv) > 3 else 0x400
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, body = None, []
for r in rows:
    if r and r[0] == 'Kernel Name':
        if body: break
        name = r[1]
    elif r and r[0] == 'Address': hdr = r
    elif hdr and len(r) > 10 and r[0].startswith('0x'): body.append(r)
ia, isrc, ismp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
print(name[:100])
tot, totex = sum(int(r[ismp]) for r in body), sum(int(r[iex]) for r in body)
print('samples', tot, 'warp-instr executed', totex)
base = int(body[0][ia], 16)
b = collections.OrderedDict()
for r in body:
    k = (int(r[ia], 16) - base) // bucket
    b.setdefault(k, [0, 0]); b[k][0] += int(r[ismp]); b[k][1] += int(r[iex])
for k, v in b.items():
    if v[0] > tot / 200: print(hex(k * bucket), 'samples %5.1f%%' % (100 * v[0] / tot), 'instr %5.1f%%' % (100 * v[1] / totex))
for r in sorted(body, key=lambda r: -int(r[ismp]))[:16]:
    print(r[ismp].rjust(7), r[iex].rjust(10), hex(int(r[ia], 16) - base), r[isrc][:90])
