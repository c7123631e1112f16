#!/usr/bin/env python
"""Per-kernel bandwidth sweep on one GPU.

    python tools/sweep.py [--nqubits 30] [--dtype complex128] [--out gpurun_out/sweep.json]

For every (k, target placement, route) it times one gate pass with CUDA events (1 warm-up +
`--reps` timed launches) and reports algorithmic GB/s = 2 * 2^n * A / 2^c / t.
"""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from qibojit_b200.backends.b200 import B200Backend

    ap = argparse.ArgumentParser()
    ap.add_argument("--nqubits", type=int, default=30)
    ap.add_argument("--dtype", default="complex128")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--kmax", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--special", action="store_true")
    args = ap.parse_args()

    b = B200Backend()
    b.set_dtype(args.dtype)
    n = args.nqubits
    amp = 16 if args.dtype == "complex128" else 8
    state = b.zero_state(n, dtype=args.dtype)
    state.fill_(1.0 / np.sqrt(2.0 ** n))
    rng = np.random.default_rng(0)
    rows = []

    def timeit(fn, alg_bytes, label):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        rows.append(dict(label=label, ms=ms, gbs=gbs))
        print(f"{label:58s} {ms:9.3f} ms {gbs:9.1f} GB/s", flush=True)

    def unitary(k):
        z = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
        q, _ = np.linalg.qr(z)
        return q.astype(args.dtype)

    full = 2.0 * amp * 2.0 ** n
    placements = {
        "low": lambda k: list(range(k)),
        "low+1": lambda k: list(range(1, k + 1)),
        "bits3..": lambda k: list(range(3, 3 + k)),
        "mid": lambda k: list(range(12, 12 + k)),
        "high": lambda k: list(range(n - k, n)),
        "spread": lambda k: sorted(int(v) for v in np.linspace(0, n - 1, k + 2)[1:-1]) if k > 1 else [n // 2],
        "lo-hi": lambda k: ([0] + list(range(n - k + 1, n))) if k > 1 else [0],
    }
    for k in range(1, args.kmax + 1):
        m = unitary(k)
        for pname, pf in placements.items():
            bits = pf(k)
            if len(set(bits)) != k:
                continue
            targets = [n - 1 - bpos for bpos in bits][::-1]  # qibo qubit numbers, MSB-first order
            for route, rname in ((1, "direct"), (2, "tile")):
                b.set_route(route)
                timeit(lambda: b._multi_qubit_base(state, n, targets, m, None), full,
                       f"{args.dtype} k={k} {pname:8s} bits={bits} {rname}")
        # controlled variants (one control on the top bit / on bit 0)
        for cname, cbit in (("ctrl-top", n - 1), ("ctrl-0", 0)):
            bits = list(range(8, 8 + k))
            targets = [n - 1 - bpos for bpos in bits][::-1]
            q = np.array(sorted(bits + [cbit]), dtype=np.int32)
            for route, rname in ((1, "direct"), (2, "tile")):
                b.set_route(route)
                timeit(lambda: b._multi_qubit_base(state, n, targets, m, q), full / 2,
                       f"{args.dtype} k={k} {cname:8s} bits={bits} {rname}")
    b.set_route(0)
    if args.special:
        for bit in (0, 1, 4, 12, n - 1):
            t = n - 1 - bit
            timeit(lambda: b._one_qubit_base(state, n, t, "apply_x", None, None), full, f"x bit={bit}")
            timeit(lambda: b._one_qubit_base(state, n, t, "apply_y", None, None), full, f"y bit={bit}")
            timeit(lambda: b._one_qubit_base(state, n, t, "apply_z", None, None), full / 2, f"z bit={bit}")
            ph = np.exp(0.3j)
            timeit(lambda: b._one_qubit_base(state, n, t, "apply_z_pow", ph, None), full / 2, f"zpow bit={bit}")
            for cb in (0, 5, n - 2):
                if cb == bit:
                    continue
                q = np.array(sorted([bit, cb]), dtype=np.int32)
                timeit(lambda: b._one_qubit_base(state, n, t, "apply_z_pow", ph, q), full / 4,
                       f"czpow bit={bit} ctrl={cb}")
                timeit(lambda: b._one_qubit_base(state, n, t, "apply_x", None, q), full / 2,
                       f"cnot bit={bit} ctrl={cb}")
        for b1, b2 in ((0, 1), (0, n - 1), (5, 17), (n - 2, n - 1)):
            timeit(lambda: b._two_qubit_base(state, n, n - 1 - b2, n - 1 - b1, "apply_swap", None, None),
                   full / 2, f"swap bits=({b1},{b2})")
            g = np.array([0.5, 0.5j, 0.5j, 0.5, np.exp(0.2j)], dtype=args.dtype)
            timeit(lambda: b._two_qubit_base(state, n, n - 1 - b2, n - 1 - b1, "apply_fsim", g, None),
                   full * 0.75, f"fsim bits=({b1},{b2})")
        timeit(lambda: b.zero_state(n, dtype=args.dtype), full / 2, "zero_state (alloc + init)")
        from qibojit_b200 import _capi
        timeit(lambda: _capi.check(b._lib.qj_initial_state(b._handle(), state.data_ptr(), b._tag(state), n)),
               full / 2, "initial_state kernel")
        state.fill_(1.0 / np.sqrt(2.0 ** n))
        timeit(lambda: b.calculate_probabilities(state, [0, 1, 2], n), full / 2, "probabilities 3 high qubits")
        timeit(lambda: b.calculate_probabilities(state, [n - 1, n - 2], n), full / 2, "probabilities 2 low qubits")
        timeit(lambda: b.calculate_norm(state), full / 2, "norm")
        if n <= 31:
            timeit(lambda: b.calculate_probabilities(state, list(range(n)), n), full * 0.75, "probabilities all qubits")
        timeit(lambda: b.collapse_state(state, [0], 0, n, True), full / 2 * 1.5, "collapse 1 qubit + normalise")
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(dict(nqubits=n, dtype=args.dtype, rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
