#!/usr/bin/env python
"""Condense an ncu report (no GPU needed) into the files committed under profiles/:

    python tools/ncu_summary.py report.ncu-rep profiles/NAME [--traffic-key qft-30-complex128]
                                [--kernel k_passId] [--units warp_tiles_per_launch]

  profiles/NAME_summary.csv    one row per metric, one column per captured launch: duration, DRAM bytes
                               (read / write), issue / pipe utilisation, occupancy, stall ratios
  profiles/NAME_sassmix.txt    SASS instruction mix of the first launch (executed warp instructions
                               and stall samples per opcode)
  profiles/NAME_lines.txt      cost per source line (tools/ncu_lines.py) of the first launch
  profiles/round2_traffic.json     with --traffic-key: dram bytes per launch (mean over the captured
                               launches) for bench.py's roofline.traffic
"""

import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
    "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_uniform.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--traffic-key", default="")
    ap.add_argument("--kernel", default="k_passId")
    ap.add_argument("--units", type=float, default=0.0)
    ap.add_argument("--lib", default=os.path.join(ROOT, "qibojit_b200", "lib", "libqibojit_b200.so"))
    args = ap.parse_args()

    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(args.out + "_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i} ({r[hdr.index('Kernel Name')][:40]})" for i, r in enumerate(data)])
        for m in METRICS + [h for h in hdr if "issue_stalled" in h and h.endswith("per_warp_active.pct")]:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in data])
    print("wrote", args.out + "_summary.csv")

    src = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
    if starts and "Instructions Executed" in srows[starts[0] + 1]:     # (metrics-only captures have no source page)
        h2 = srows[starts[0] + 1]
        end = starts[1] if len(starts) > 1 else len(srows)
        body = [r for r in srows[starts[0] + 2:end] if len(r) == len(h2)]
        iS, iN, iSm = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
        byop, samp = collections.Counter(), collections.Counter()
        for r in body:
            toks = [o for o in r[iS].split() if not o.startswith("@")]
            op = toks[0].split(".")[0] if toks else "?"
            byop[op] += int(r[iN])
            samp[op] += int(r[iSm])
        tot, tots = sum(byop.values()), sum(samp.values())
        with open(args.out + "_sassmix.txt", "w") as f:
            f.write(f"# SASS instruction mix of the first captured launch of {os.path.basename(args.report)}\n")
            f.write(f"# static SASS instructions {len(body)}, executed warp instructions {tot}, stall samples {tots}\n")
            for op, n in byop.most_common(40):
                f.write(f"{op:12s} {n:14d} {100 * n / tot:5.1f}%   samples {100 * samp[op] / max(1, tots):5.1f}%\n")
        print("wrote", args.out + "_sassmix.txt")
    if not (starts and "Instructions Executed" in srows[starts[0] + 1]):
        args.kernel = ""
    lines = None if not args.kernel else subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), args.report, args.kernel,
                            "--launch", "0", "--top", "45", "--lib", args.lib] +
                           (["--units", str(args.units)] if args.units else []), capture_output=True, text=True)
    if args.kernel:
        with open(args.out + "_lines.txt", "w") as f:
            f.write(lines.stdout + lines.stderr)
        print("wrote", args.out + "_lines.txt")

    if args.traffic_key:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        per = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in data]
        path = os.path.join(ROOT, "profiles", "round2_traffic.json")
        try:
            store = json.load(open(path))
        except Exception:
            store = {}
        store[args.traffic_key] = {
            "bytes_per_launch": sum(per) / len(per), "launches": len(per), "per_launch": per,
            "source": f"profiles/{os.path.basename(args.out)}_summary.csv (ncu: dram__bytes_read.sum + dram__bytes_write.sum, "
                      f"mean of {len(per)} k_pass launches of this workload)"}
        json.dump(store, open(path, "w"), indent=1, sort_keys=True)
        print("updated", path)


if __name__ == "__main__":
    main()
