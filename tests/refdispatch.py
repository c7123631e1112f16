"""qibo-free restatement of the reference backend's kernel dispatch (test infrastructure).

Follows /root/reference/src/qibojit/backends/cpu.py:565-569 (`_create_qubits_tensor`),
:606-616 (`_one_qubit_base`), :618-635 (`_two_qubit_base`), :581-604 (`_multi_qubit_base`)
and :541-563 (`_collapse_statevector`).  ``K`` is a kernel module with the reference's
function names: either the reference's own ``qibojit.custom_operators.gates`` (golden
generator) or ``oracle.oracle`` (parity tests).
"""

import numpy as np


def qubits_tensor(nqubits, targets, controls=()):
    qubits = [nqubits - q - 1 for q in controls]
    qubits.extend(nqubits - q - 1 for q in targets)
    return np.array(sorted(qubits), dtype=np.int32)


def one_qubit_base(K, state, nqubits, target, kernel, gate, qubits):
    ncontrols = len(qubits) - 1 if qubits is not None else 0
    m = nqubits - target - 1
    nstates = 1 << (nqubits - ncontrols - 1)
    if ncontrols:
        return getattr(K, f"multicontrol_{kernel}_kernel")(state, gate, qubits, nstates, m)
    return getattr(K, f"{kernel}_kernel")(state, gate, nstates, m)


def two_qubit_base(K, state, nqubits, target1, target2, kernel, gate, qubits):
    ncontrols = len(qubits) - 2 if qubits is not None else 0
    if target1 > target2:
        swap_targets = True
        m1, m2 = nqubits - target1 - 1, nqubits - target2 - 1
    else:
        swap_targets = False
        m1, m2 = nqubits - target2 - 1, nqubits - target1 - 1
    nstates = 1 << (nqubits - 2 - ncontrols)
    if ncontrols:
        return getattr(K, f"multicontrol_{kernel}_kernel")(
            state, gate, qubits, nstates, m1, m2, swap_targets)
    return getattr(K, f"{kernel}_kernel")(state, gate, nstates, m1, m2, swap_targets)


def multi_qubit_base(K, state, nqubits, targets, gate, qubits):
    if qubits is None:
        qubits = np.array(sorted(nqubits - q - 1 for q in targets), dtype=np.int32)
    nstates = 1 << (nqubits - len(qubits))
    tmasks = np.array([1 << (nqubits - t - 1) for t in targets[::-1]], dtype=np.int64)
    names = {3: "apply_three_qubit_gate_kernel", 4: "apply_four_qubit_gate_kernel",
             5: "apply_five_qubit_gate_kernel"}
    kernel = getattr(K, names.get(len(targets), "apply_multi_qubit_gate_kernel"))
    return kernel(state, gate, qubits, nstates, tmasks)


def collapse(O, state, qubits, shot, nqubits, normalize=True):
    q = np.array([nqubits - q - 1 for q in reversed(qubits)], dtype=np.int32)
    if normalize:
        return O.collapse_state_normalized(state, q, int(shot), nqubits)
    return O.collapse_state(state, q, int(shot), nqubits)


def random_state(nqubits, dtype, seed):
    rng = np.random.default_rng(seed)
    n = 1 << nqubits
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x /= np.linalg.norm(x)
    return x.astype(dtype)


def random_matrix(dim, dtype, seed):
    """Non-unitary random complex matrix, like tests/utils.py:10-12 of the reference."""
    rng = np.random.default_rng(seed + 7919)
    return (rng.random((dim, dim)) + 1j * rng.random((dim, dim))).astype(dtype)


def einsum_apply(state, matrix, targets, controls, nqubits):
    """Independent known-answer oracle (plays the role qibo's NumpyBackend has in the
    reference's tests): reshape to n binary axes, contract the target axes."""
    k = len(targets)
    psi = np.array(state, dtype=np.complex128).reshape(nqubits * (2,))
    mat = np.asarray(matrix, dtype=np.complex128).reshape(2 * k * (2,))
    sl = [slice(None)] * nqubits
    for c in controls:
        sl[c] = 1
    sub = psi[tuple(sl)]
    # axes of `sub` that correspond to the targets
    free = [q for q in range(nqubits) if q not in controls]
    axes = [free.index(t) for t in targets]
    out = np.tensordot(mat, sub, axes=(list(range(k, 2 * k)), axes))
    out = np.moveaxis(out, list(range(k)), axes)
    psi[tuple(sl)] = out
    return psi.reshape(-1)


def random_unitary(dim, seed):
    rng = np.random.default_rng(seed + 104729)
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def reference_run(state, gate_list, nqubits):
    """Gate-by-gate einsum application of gate objects (complex128)."""
    from qibojit_b200 import fusion
    from qibojit_b200.matrices import CustomMatrices

    mats = CustomMatrices("complex128")
    ref = np.array(state, dtype=np.complex128)
    for g in gate_list:
        if g.name == "fanout":
            for tq in g.target_qubits:
                ref = einsum_apply(ref, mats.X, [tq], [g.control_qubits[0]], nqubits)
        else:
            ref = einsum_apply(ref, fusion.target_only_matrix(g, mats), list(g.target_qubits),
                               list(g.control_qubits), nqubits)
    return ref
