"""Generate the committed golden vectors from the REFERENCE's own numba kernels.

Run in the build container only (needs /root/reference and numba):

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py

It imports ``qibojit.custom_operators.{gates,ops}`` straight from /root/reference through a
package stub (``import qibojit`` itself needs qibo, which is not installed), drives them with
the dispatch restated in tests/refdispatch.py on seeded inputs, and stores the outputs in
tests/golden/*.npz.  Inputs are re-derived from the seeds by the tests, so only outputs are
stored.  Large (20-qubit) outputs are stored as a strided subsample plus a projection on a
seeded random vector.
"""

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")

pkg = types.ModuleType("qibojit")
pkg.__path__ = ["/root/reference/src/qibojit"]
sys.modules["qibojit"] = pkg
import qibojit.custom_operators.gates as G  # noqa: E402
import qibojit.custom_operators.ops as O  # noqa: E402

from tests import cases  # noqa: E402
from tests import refdispatch as R  # noqa: E402


def case_key(kind, dtype, nq, targets, controls, extra=""):
    t = "_".join(map(str, targets)) if isinstance(targets, (list, tuple)) else str(targets)
    c = "_".join(map(str, controls))
    return f"{kind}|{dtype}|n{nq}|t{t}|c{c}|{extra}"


def subsample(out, seed):
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal(out.size) + 1j * rng.standard_normal(out.size))
    return np.concatenate([out[::4099].astype(np.complex128), [np.vdot(w, out.astype(np.complex128))]])


def gates_golden():
    store = {}
    for dtype in cases.DTYPES:
        seed = 0
        for nq, t, c in cases.ONE_QUBIT:
            seed += 1
            st = R.random_state(nq, dtype, seed)
            m = R.random_matrix(2, dtype, seed)
            q = R.qubits_tensor(nq, [t], c)
            out = R.one_qubit_base(G, st, nq, t, "apply_gate", m, q if c else None)
            store[case_key("gate", dtype, nq, t, c)] = out
        for name in ("x", "y", "z"):
            for nq, t, c in cases.PAULI:
                seed += 1
                st = R.random_state(nq, dtype, seed)
                q = R.qubits_tensor(nq, [t], c)
                out = R.one_qubit_base(G, st, nq, t, f"apply_{name}", None, q if c else None)
                store[case_key(name, dtype, nq, t, c)] = out
        for nq, t, c in cases.ZPOW:
            seed += 1
            st = R.random_state(nq, dtype, seed)
            phase = np.exp(1j * 0.1234 * seed).astype(dtype)
            q = R.qubits_tensor(nq, [t], c)
            out = R.one_qubit_base(G, st, nq, t, "apply_z_pow", phase, q if c else None)
            store[case_key("zpow", dtype, nq, t, c)] = out
        for nq, t, c in cases.TWO_QUBIT:
            seed += 1
            st = R.random_state(nq, dtype, seed)
            m = R.random_matrix(4, dtype, seed)
            q = R.qubits_tensor(nq, t, c)
            out = R.two_qubit_base(G, st, nq, t[0], t[1], "apply_two_qubit_gate", m,
                                   q if c else None)
            store[case_key("two", dtype, nq, t, c)] = out
        for nq, t, c in cases.SWAP:
            seed += 1
            st = R.random_state(nq, dtype, seed)
            q = R.qubits_tensor(nq, t, c)
            out = R.two_qubit_base(G, st, nq, t[0], t[1], "apply_swap", None, q if c else None)
            store[case_key("swap", dtype, nq, t, c)] = out
        for nq, t, c in cases.FSIM:
            seed += 1
            st = R.random_state(nq, dtype, seed)
            m = R.random_matrix(3, dtype, seed).ravel()[:5].copy()
            q = R.qubits_tensor(nq, t, c)
            out = R.two_qubit_base(G, st, nq, t[0], t[1], "apply_fsim", m, q if c else None)
            store[case_key("fsim", dtype, nq, t, c)] = out
        for nq, t, c in cases.MULTI_QUBIT:
            seed += 1
            st = R.random_state(nq, dtype, seed)
            m = R.random_matrix(1 << len(t), dtype, seed)
            q = R.qubits_tensor(nq, t, c)
            out = R.multi_qubit_base(G, st, nq, t, m, q)
            store[case_key("multi", dtype, nq, t, c)] = out
        for nq, t, c in cases.MULTI_QUBIT_LARGE:
            seed += 1
            st = R.random_state(nq, dtype, seed)
            m = R.random_matrix(1 << len(t), dtype, seed)
            q = R.qubits_tensor(nq, t, c)
            out = R.multi_qubit_base(G, st, nq, t, m, q)
            store[case_key("multilarge", dtype, nq, t, c)] = subsample(out, seed)
    np.savez_compressed(os.path.join(HERE, "gates_golden.npz"), **store)
    return len(store)


def ops_golden():
    store = {}
    for dtype in cases.DTYPES:
        seed = 1000
        st = np.empty(1 << 7, dtype=dtype)
        store[f"init|{dtype}"] = O.initial_state_vector(st)
        for nq, meas, res in cases.COLLAPSE:
            for normalize in (True, False):
                seed += 1
                st = R.random_state(nq, dtype, seed)
                shot = int("".join(map(str, res)), 2)
                out = R.collapse(O, st, meas, shot, nq, normalize)
                if nq > 12:
                    out = subsample(out, seed)
                store[case_key("collapse", dtype, nq, meas, res, f"norm{int(normalize)}")] = out
        # swap_pieces (ops.py:131-137): two pieces of nlocal qubits
        for nlocal, new_global in [(3, 0), (3, 2), (5, 1), (9, 0), (9, 4), (9, 8)]:
            seed += 1
            full = R.random_state(nlocal + 1, dtype, seed)
            p0, p1 = full[: 1 << nlocal].copy(), full[1 << nlocal:].copy()
            O.swap_pieces(p0, p1, new_global, nlocal)
            store[f"swap_pieces|{dtype}|l{nlocal}|g{new_global}"] = np.concatenate([p0, p1])
        # transpose_state (ops.py:112-124)
        rng = np.random.default_rng(77)
        for nq, ndev in [(3, 2), (5, 4), (8, 8), (10, 2)]:
            seed += 1
            order = [int(v) for v in rng.permutation(nq)]
            full = R.random_state(nq, dtype, seed)
            pieces = [p.copy() for p in full.reshape(ndev, -1)]
            from numba.typed import List as NList
            pl = NList()
            for p in pieces:
                pl.append(p)
            out = O.transpose_state(pl, np.zeros_like(full), nq, np.array(order))
            store[f"transpose|{dtype}|n{nq}|d{ndev}|o{'_'.join(map(str, order))}"] = out
    # measure_frequencies (ops.py:86-108): RNG stream pinned for several shapes
    for realtype in ("float32", "float64"):
        for nq, nshots, seedv, nthreads in [(4, 1000, 1234, 4), (4, 1000, 1234, 1), (6, 5000, 42, 3),
                                            (10, 20000, 7, 8), (3, 777, 99999999, 5)]:
            rng = np.random.default_rng(nq * 131 + nshots)
            probs = rng.random(1 << nq)
            if nq == 4:
                probs = np.ones(16)
            probs = (probs / probs.sum()).astype(realtype)
            freq = np.zeros(1 << nq, dtype=np.int64)
            out = O.measure_frequencies(freq, probs, nshots, nq, seedv, nthreads)
            store[f"freq|{realtype}|n{nq}|s{nshots}|seed{seedv}|t{nthreads}"] = out
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **store)
    return len(store)


if __name__ == "__main__":
    print("gate cases:", gates_golden())
    print("ops cases:", ops_golden())
