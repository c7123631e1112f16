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
