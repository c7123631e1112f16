"""Helpers shared by the golden generator's consumers (key format + subsampling)."""
import numpy as np


def case_key(kind, dtype, nq, targets, controls, extra=""):
    t = "_".join(map(str, targets)) if isinstance(targets, (list, tuple)) else str(targets)
    c = "_".join(map(str, controls))
    return f"{kind}|{dtype}|n{nq}|t{t}|c{c}|{extra}"


def subsample(out, seed):
    rng = np.random.default_rng(seed)
    w = rng.standard_normal(out.size) + 1j * rng.standard_normal(out.size)
    return np.concatenate([out[::4099].astype(np.complex128),
                           [np.vdot(w, out.astype(np.complex128))]])


def iter_gate_cases(cases, R, large=True):
    """Yield (key, kind, dtype, nq, targets, controls, seed) in the generator's seed order."""
    for dtype in cases.DTYPES:
        seed = 0
        for nq, t, c in cases.ONE_QUBIT:
            seed += 1
            yield case_key("gate", dtype, nq, t, c), "gate", dtype, nq, t, c, seed
        for name in ("x", "y", "z"):
            for nq, t, c in cases.PAULI:
                seed += 1
                yield case_key(name, dtype, nq, t, c), name, dtype, nq, t, c, seed
        for nq, t, c in cases.ZPOW:
            seed += 1
            yield case_key("zpow", dtype, nq, t, c), "zpow", dtype, nq, t, c, seed
        for nq, t, c in cases.TWO_QUBIT:
            seed += 1
            yield case_key("two", dtype, nq, t, c), "two", dtype, nq, t, c, seed
        for nq, t, c in cases.SWAP:
            seed += 1
            yield case_key("swap", dtype, nq, t, c), "swap", dtype, nq, t, c, seed
        for nq, t, c in cases.FSIM:
            seed += 1
            yield case_key("fsim", dtype, nq, t, c), "fsim", dtype, nq, t, c, seed
        for nq, t, c in cases.MULTI_QUBIT:
            seed += 1
            yield case_key("multi", dtype, nq, t, c), "multi", dtype, nq, t, c, seed
        for nq, t, c in cases.MULTI_QUBIT_LARGE:
            seed += 1
            if large:
                yield case_key("multilarge", dtype, nq, t, c), "multilarge", dtype, nq, t, c, seed


def run_gate_case(B, R, kind, dtype, nq, t, c, seed, state=None):
    """Apply one table case through the dispatch `B` = (one, two, multi) callables.

    Returns the output state.  `state` lets the caller supply a device copy of the
    seeded input; by default the numpy input is used."""
    one, two, multi = B
    st = R.random_state(nq, dtype, seed) if state is None else state
    if kind == "gate":
        q = R.qubits_tensor(nq, [t], c)
        return one(st, nq, t, "apply_gate", R.random_matrix(2, dtype, seed), q if c else None)
    if kind in ("x", "y", "z"):
        q = R.qubits_tensor(nq, [t], c)
        return one(st, nq, t, f"apply_{kind}", None, q if c else None)
    if kind == "zpow":
        q = R.qubits_tensor(nq, [t], c)
        phase = np.exp(1j * 0.1234 * seed).astype(dtype)
        return one(st, nq, t, "apply_z_pow", phase, q if c else None)
    if kind == "two":
        q = R.qubits_tensor(nq, t, c)
        return two(st, nq, t[0], t[1], "apply_two_qubit_gate", R.random_matrix(4, dtype, seed),
                   q if c else None)
    if kind == "swap":
        q = R.qubits_tensor(nq, t, c)
        return two(st, nq, t[0], t[1], "apply_swap", None, q if c else None)
    if kind == "fsim":
        q = R.qubits_tensor(nq, t, c)
        m = R.random_matrix(3, dtype, seed).ravel()[:5].copy()
        return two(st, nq, t[0], t[1], "apply_fsim", m, q if c else None)
    if kind in ("multi", "multilarge"):
        q = R.qubits_tensor(nq, t, c)
        return multi(st, nq, t, R.random_matrix(1 << len(t), dtype, seed), q)
    raise ValueError(kind)
