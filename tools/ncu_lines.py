#!/usr/bin/env python
"""Per-source-line cost of a kernel from an ncu report (no GPU needed).

    python tools/ncu_lines.py report.ncu-rep k_passId [--launch 0] [--top 40] [--units N]

Joins `ncu --page source --print-source sass` (stall samples and instructions executed per SASS
instruction) with `nvdisasm -g` of the in-tree library (SASS instruction -> source line, following
"inlined at" chains up to the kernel's own file) and prints, per source line of the kernel file,
the share of samples and the warp instructions executed (divided by --units when given, e.g. the
number of warp-tiles, so the figure reads "instructions per warp per tile")."""

import argparse
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(lib, kernel_substr, src_name):
    """[(opcode text, line)] for the first kernel whose mangled name contains `kernel_substr`."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        inside, line = False, 0
        for ln in dis.splitlines():
            if ln.startswith(".text."):
                if inside and out:
                    return out
                inside = kernel_substr in ln
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cands = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
                own = [l for f_, l in cands if f_.endswith(src_name)]
                line = own[0] if own else -1
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((m.group(2).strip(), line))
        if out:
            return out
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--units", type=float, default=0.0)
    ap.add_argument("--lib", default=os.path.join(ROOT, "qibojit_b200", "lib", "libqibojit_b200.so"))
    ap.add_argument("--src", default="pass_kernels.cu")
    args = ap.parse_args()

    raw = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    s = starts[args.launch]
    e = starts[args.launch + 1] if args.launch + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    body = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
    ci, cs, cx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    lines = sass_lines(args.lib, args.kernel, args.src)
    if len(lines) != len(body):
        print(f"warning: {len(body)} instructions in the report, {len(lines)} in the library", file=sys.stderr)
    per = {}
    tot_s = tot_x = 0
    for k, r in enumerate(body):
        line = lines[k][1] if k < len(lines) else -1
        d = per.setdefault(line, [0, 0, 0])
        d[0] += int(r[cs]); d[1] += int(r[cx]); d[2] += 1
        tot_s += int(r[cs]); tot_x += int(r[cx])
    src = open(os.path.join(ROOT, "qibojit_b200", "csrc", args.src)).read().splitlines()
    div = args.units or 1.0
    print(f"total: {tot_s} samples, {tot_x / div:.1f} warp instructions" + (" per unit" if args.units else ""))
    for line, (smp, ex, n) in sorted(per.items(), key=lambda kv: -kv[1][0])[:args.top]:
        text = src[line - 1].strip()[:90] if 0 < line <= len(src) else "?"
        print(f"{100 * smp / tot_s:5.1f}% smp {ex / div:9.1f} inst {n:5d} sass  L{line}: {text}")


if __name__ == "__main__":
    main()
