#!/usr/bin/env python
"""Where a kernel spends its time, from an ncu report:

    ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv
    python tools/ncu_blocks.py x.csv [min_share]

Prints the stall mix over all samples and, for every run of instructions with the same execution
count (a basic block) that holds at least `min_share` (default 1.5 %) of the samples: its share of
instructions and samples, its opcode mix and its hottest instruction with the dominant stall."""

import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
iS, iN, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows if len(r) >= len(hdr) and r[0].startswith("0x")]
tot = sum(int(r[iN]) for r in body)
tots = sum(int(r[iSm]) for r in body)
print("total warp instructions", tot, "samples", tots)
st = collections.Counter()
for r in body:
    for c in stall_cols:
        st[hdr[c]] += int(r[c])
print([(k, round(100 * v / max(1, tots), 1)) for k, v in st.most_common(9)])
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.015


def opcode(src):
    toks = src.split()
    return toks[1 if toks[0].startswith("@") else 0].split(".")[0]


i = 0
while i < len(body):
    j = i
    n = int(body[i][iN])
    while j < len(body) and int(body[j][iN]) == n:
        j += 1
    smp = sum(int(r[iSm]) for r in body[i:j])
    if n > 0 and smp > thr * tots:
        ops = collections.Counter(opcode(body[k][iS]) for k in range(i, j))
        top = max(range(i, j), key=lambda k: int(body[k][iSm]))
        why = sorted(((int(body[top][c]), hdr[c]) for c in stall_cols), reverse=True)[0]
        print(f"{i:6d}-{j:6d} len {j - i:4d} x {n:11d} = {100 * (j - i) * n / tot:5.1f}% inst, "
              f"{100 * smp / tots:5.1f}% samples  {ops.most_common(4)}  hot: {body[top][iS][:46]} "
              f"({int(body[top][iSm])} {why[1]})")
    i = j
