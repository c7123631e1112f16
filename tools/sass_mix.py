#!/usr/bin/env python
"""Instruction mix of a kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
iS, iN, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = 0
byop = collections.Counter()
samp = collections.Counter()
totsamp = 0
for r in rows:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    n, s = int(r[iN]), int(r[iSm])
    toks = [o for o in r[iS].split() if not o.startswith("@")]
    op = toks[0].split(".")[0] if toks else "?"
    byop[op] += n
    samp[op] += s
    tot += n
    totsamp += s
print("total warp inst", tot, "samples", totsamp)
for op, n in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:12s} {n:14d} {100*n/tot:5.1f}%   samples {100*samp[op]/max(1,totsamp):5.1f}%")
