#!/usr/bin/env python
"""Per-pass timing of the multi-gate pass program of a benchmark circuit on one GPU.

    python tools/prog_bench.py --workload qft --nqubits 30 [--dtype complex128] [--tile-bits 12]
                               [--run-bits 5] [--diag-bits 10] [--reps 3]
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from qibojit_b200 import _capi, circuits, planner
    from qibojit_b200.backends.b200 import B200Backend

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="qft")
    ap.add_argument("--nqubits", type=int, default=30)
    ap.add_argument("--dtype", default="complex128")
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--run-bits", type=int, default=0)
    ap.add_argument("--diag-bits", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--keep-swaps", action="store_true", help="do not compile for the |0...0> input (SWAP gates stay)")
    ap.add_argument("--no-pair-blocks", action="store_true", help="keep one- and two-qubit gates apart (no 4x4 blocks)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    b = B200Backend()
    b.set_dtype(args.dtype)
    n = args.nqubits
    build = {"qft": circuits.qft, "variational": circuits.variational, "supremacy": circuits.supremacy,
             "qv": circuits.quantum_volume}[args.workload]
    c = build(n)
    t0 = time.perf_counter()
    prog = planner.Program(b, c.queue, n, dtype=args.dtype, tile_bits=args.tile_bits or None,
                           run_bits=args.run_bits or None, max_diag_bits=args.diag_bits,
                           zero_state=not args.keep_swaps, pair_blocks=not args.no_pair_blocks)
    t_plan = time.perf_counter() - t0
    stats = prog.stats()
    state = b.zero_state(n)
    amp = 16 if args.dtype == "complex128" else 8
    full = 2.0 * amp * 2.0 ** n
    launches = prog.launches()
    rows = []
    for rep in range(args.reps + 1):
        _capi.check(b._lib.qj_initial_state(b._handle(), state.data_ptr(), b._tag(state), n))
        torch.cuda.synchronize()
        evs = []
        s0 = torch.cuda.Event(enable_timing=True); s0.record()
        for (h, i) in launches:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _capi.check(b._lib.qj_program_run_launch(b._handle(), h, state.data_ptr(), i))
            e1.record()
            evs.append((e0, e1))
        s1 = torch.cuda.Event(enable_timing=True); s1.record()
        torch.cuda.synchronize()
        if rep:
            rows.append(([e0.elapsed_time(e1) for e0, e1 in evs], s0.elapsed_time(s1)))
    per = np.median(np.array([r[0] for r in rows]), axis=0)
    total = float(np.median([r[1] for r in rows]))
    norm = b.calculate_norm(state)
    print(f"{args.workload}-{n} {args.dtype} T={prog.tile_bits} r={prog.run_bits}: {len(c.queue)} gates -> "
          f"{stats['passes']} passes, {stats['launches']} launches, {stats['rounds']} rounds, "
          f"{stats['micro_ops']} micro-ops, {stats['raw_gates']} raw; plan+compile {t_plan*1e3:.0f} ms")
    for (lb, rounds), ms in zip(prog.passes, per):
        ops = [o for _, ro in rounds for o in ro]
        nd = sum(1 for o in ops if o.kind == "diag")
        print(f"  pass hi={lb[prog.run_bits:]} rounds={len(rounds)} ops={len(ops)} (dense {len(ops)-nd}, diag {nd}): "
              f"{ms:8.3f} ms {full/ms/1e6:8.1f} GB/s")
    print(f"  total {total:.3f} ms -> {len(c.queue)/total*1e3:.1f} gates/s; norm={norm:.12f}")
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(dict(workload=args.workload, n=n, dtype=args.dtype, T=prog.tile_bits,
                                    r=prog.run_bits, stats=stats, per_launch_ms=[float(x) for x in per],
                                    total_ms=total, gates=len(c.queue), norm=norm)) + "\n")


if __name__ == "__main__":
    main()
