"""TEST INFRASTRUCTURE: seeded random circuits over the gate set the planner lowers."""

import numpy as np

from qibojit_b200 import gates
from tests import refdispatch as R


def random_circuit_gates(n, ngates, seed, dtype="complex128", with_raw=True):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(ngates):
        kind = int(rng.integers(0, 14))
        qs = [int(v) for v in rng.permutation(n)[:4]]
        if kind == 0:
            out.append(gates.H(qs[0]))
        elif kind == 1:
            out.append(gates.RY(qs[0], float(rng.uniform(0, 6))))
        elif kind == 2:
            out.append(gates.CZ(qs[0], qs[1]))
        elif kind == 3:
            out.append(gates.CU1(qs[0], qs[1], float(rng.uniform(0, 6))))
        elif kind == 4:
            out.append(gates.CNOT(qs[0], qs[1]))
        elif kind == 5:
            out.append(gates.RZ(qs[0], float(rng.uniform(0, 6))))
        elif kind == 6:
            out.append(gates.Unitary(R.random_unitary(4, i + 7 * seed), qs[0], qs[1]))
        elif kind == 7:
            out.append(gates.SWAP(qs[0], qs[1]))
        elif kind == 8:
            out.append(gates.fSim(qs[0], qs[1], float(rng.uniform(0, 3)), float(rng.uniform(0, 3))))
        elif kind == 9:
            out.append(gates.TOFFOLI(qs[0], qs[1], qs[2]))
        elif kind == 10:
            out.append(gates.U3(qs[0], 0.3, 0.9, 1.7).controlled_by(qs[1], qs[2]))
        elif kind == 11:
            out.append(gates.CCZ(qs[0], qs[1], qs[2]))
        elif kind == 12:
            out.append(gates.Unitary(R.random_unitary(4, i + 11 * seed), qs[0], qs[1]).controlled_by(qs[2]))
        elif kind == 13 and with_raw:
            out.append(gates.Unitary(R.random_unitary(8, i + 13 * seed), qs[0], qs[1], qs[2]))
        else:
            out.append(gates.T(qs[0]))
    return out


_ONE = ["H", "X", "Y", "Z", "S", "SDG", "T", "TDG", "SX", "SXDG", "I"]
_ONEP = ["RX", "RY", "RZ", "U1", "GPI", "GPI2"]
_TWO_C = ["CNOT", "CY", "CZ", "CH", "CSX", "CSXDG"]
_TWO_CP = ["CRX", "CRY", "CRZ", "CU1"]
_TWO_T = ["SWAP", "iSWAP", "SiSWAP", "FSWAP"]
_TWO_TP = ["RXX", "RYY", "RZZ"]


def random_gate_any(n, rng):
    """One gate drawn from the whole gate table of qibojit_b200.gates (every matrix kind the
    pass kernel distinguishes: real, axis-aligned, complex, permutations, phases), sometimes with
    up to two extra controls."""
    k = int(rng.integers(0, 12))
    qs = [int(v) for v in rng.permutation(n)[:4]]
    th = float(rng.uniform(0, 6.28))
    pick = lambda names: getattr(gates, names[int(rng.integers(len(names)))])  # noqa: E731
    if k == 0:
        g = pick(_ONE)(qs[0])
    elif k == 1:
        g = pick(_ONEP)(qs[0], th)
    elif k == 2:
        g = pick(_TWO_C)(qs[0], qs[1])
    elif k == 3:
        g = pick(_TWO_CP)(qs[0], qs[1], th)
    elif k == 4:
        g = pick(_TWO_T)(qs[0], qs[1])
    elif k == 5:
        g = pick(_TWO_TP)(qs[0], qs[1], th)
    elif k == 6:
        g = gates.U3(qs[0], th, 0.3, 1.1)
    elif k == 7:
        g = gates.U2(qs[0], th, 0.4)
    elif k == 8:
        g = gates.fSim(qs[0], qs[1], th, 0.7)
    elif k == 9:
        g = gates.GeneralizedfSim(qs[0], qs[1], R.random_unitary(2, int(rng.integers(1 << 20))), th)
    elif k == 10:
        g = gates.FanOut(qs[0], qs[1], qs[2])
    else:
        g = gates.Unitary(R.random_unitary(2, int(rng.integers(1 << 20))), qs[0])
    if k != 10 and rng.random() < 0.3:
        extra = [q for q in qs if q not in g.qubits][:int(rng.integers(0, 3))]
        if extra:
            g = g.controlled_by(*extra)
    return g
