"""The benchmark circuits of BASELINE.json, built from synthetic deterministic parameters.

Only the QFT is defined inside the reference (benchmarks/abstract.py:86-104); the
variational, supremacy-style and quantum-volume circuits live in the external
qibojit-benchmarks repository and are defined here as SURVEY.md section 8(d) specifies.
"""

import numpy as np

from . import gates
from .circuit import Circuit


def qft(nqubits, swaps=True):
    """H / controlled-phase ladder / final swaps, the schedule of benchmarks/abstract.py:86-104."""
    c = Circuit(nqubits)
    for i1 in range(nqubits):
        c.add(gates.H(i1))
        for i2 in range(i1 + 1, nqubits):
            c.add(gates.CU1(i2, i1, np.pi / 2 ** (i2 - i1)))
    if swaps:
        for i in range(nqubits // 2):
            c.add(gates.SWAP(i, nqubits - i - 1))
    return c


def variational(nqubits, nlayers=2, seed=123):
    """nlayers x [RY all; CZ even pairs; RY all; CZ odd pairs; CZ(0, n-1)] + final RY layer."""
    rng = np.random.default_rng(seed)
    c = Circuit(nqubits)
    for _ in range(nlayers):
        c.add(gates.RY(q, float(rng.uniform(0, 2 * np.pi))) for q in range(nqubits))
        c.add(gates.CZ(q, q + 1) for q in range(0, nqubits - 1, 2))
        c.add(gates.RY(q, float(rng.uniform(0, 2 * np.pi))) for q in range(nqubits))
        c.add(gates.CZ(q, q + 1) for q in range(1, nqubits - 1, 2))
        c.add(gates.CZ(0, nqubits - 1))
    c.add(gates.RY(q, float(rng.uniform(0, 2 * np.pi))) for q in range(nqubits))
    return c


def supremacy(nqubits, depth=8, seed=123):
    """H on all; `depth` cycles of [random sqrt-X / sqrt-Y / sqrt-W per qubit; CZ brick that
    shifts every cycle]; final H layer."""
    rng = np.random.default_rng(seed)
    c = Circuit(nqubits)
    c.add(gates.H(q) for q in range(nqubits))
    sqrt_w = (np.array([[1, -np.sqrt(1j)], [np.sqrt(-1j), 1]]) / np.sqrt(2))
    for cycle in range(depth):
        for q in range(nqubits):
            kind = int(rng.integers(0, 3))
            if kind == 0:
                c.add(gates.RX(q, np.pi / 2))
            elif kind == 1:
                c.add(gates.RY(q, np.pi / 2))
            else:
                c.add(gates.Unitary(sqrt_w, q))
        start = cycle % 2
        c.add(gates.CZ(q, q + 1) for q in range(start, nqubits - 1, 2))
    c.add(gates.H(q) for q in range(nqubits))
    return c


def _haar_unitary(dim, rng):
    z = (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))) / np.sqrt(2)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def quantum_volume(nqubits, depth=None, seed=123):
    """`depth` layers of [random qubit permutation; Haar-random SU(4) on adjacent pairs]."""
    rng = np.random.default_rng(seed)
    depth = nqubits if depth is None else depth
    c = Circuit(nqubits)
    for _ in range(depth):
        perm = rng.permutation(nqubits)
        for i in range(0, nqubits - 1, 2):
            c.add(gates.Unitary(_haar_unitary(4, rng), int(perm[i]), int(perm[i + 1])))
    return c
