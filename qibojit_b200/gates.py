"""Minimal qibo-compatible gate objects.

qibo (``qibo.gates``) is a third-party dependency of the reference that is not available
offline; these classes carry exactly the attributes the backend dispatch reads
(/root/reference/src/qibojit/backends/cpu.py:433-450, 519-539, 565-569 and SURVEY.md
appendix C): ``name``, the class name, ``target_qubits``, ``control_qubits``, ``qubits``,
``parameters``, ``init_args``, ``init_kwargs``.  A real ``qibo.gates`` object works in their
place -- the backend only duck-types.
"""

import numpy as np


class Gate:
    ntargets = 1
    nparams = 0
    parametrized = False
    diagonal = False  # diagonal in the computational basis (exchange-free on global qubits)

    def __init__(self, *args, **kwargs):
        nq = self.nqubit_args
        qubits = tuple(int(q) for q in args[:nq])
        self.init_args = list(args)
        self.init_kwargs = dict(kwargs)
        self.parameters = tuple(args[nq:]) + tuple(kwargs.values())
        ncontrols = len(qubits) - self.ntargets
        self.control_qubits = tuple(qubits[:ncontrols])
        self.target_qubits = tuple(qubits[ncontrols:])
        self.name = self.__class__.__name__.lower()
        self.is_controlled_by = False
        self._check()

    def _check(self):
        q = self.qubits
        if len(set(q)) != len(q):
            raise ValueError(f"duplicate qubits in gate {self.name}: {q}")

    @property
    def qubits(self):
        return tuple(self.control_qubits) + tuple(self.target_qubits)

    def controlled_by(self, *qubits):
        if qubits:
            self.control_qubits = tuple(self.control_qubits) + tuple(int(q) for q in qubits)
            self.is_controlled_by = True
            self._check()
        return self

    def target_matrix(self, matrices):
        """Matrix the kernel consumes (target qubits only), from a CustomMatrices table."""
        entry = getattr(matrices, self.__class__.__name__)
        if callable(entry):
            return entry(*self.parameters)
        return entry

    def apply(self, backend, state, nqubits):
        return backend.apply_gate(self, state, nqubits)

    def __repr__(self):
        return (f"{self.__class__.__name__}(targets={self.target_qubits}, "
                f"controls={self.control_qubits}, params={self.parameters})")


def _make(name, nqubit_args, ntargets=1, parametrized=False, diagonal=False):
    cls = type(name, (Gate,), {"nqubit_args": nqubit_args, "ntargets": ntargets,
                               "parametrized": parametrized, "diagonal": diagonal})
    return cls


H = _make("H", 1)
X = _make("X", 1)
Y = _make("Y", 1)
Z = _make("Z", 1, diagonal=True)
S = _make("S", 1, diagonal=True)
SDG = _make("SDG", 1, diagonal=True)
T = _make("T", 1, diagonal=True)
TDG = _make("TDG", 1, diagonal=True)
SX = _make("SX", 1)
SXDG = _make("SXDG", 1)
I = _make("I", 1, diagonal=True)
RX = _make("RX", 1, parametrized=True)
RY = _make("RY", 1, parametrized=True)
RZ = _make("RZ", 1, parametrized=True, diagonal=True)
GPI = _make("GPI", 1, parametrized=True)
GPI2 = _make("GPI2", 1, parametrized=True)
U1 = _make("U1", 1, parametrized=True, diagonal=True)
U2 = _make("U2", 1, parametrized=True)
U3 = _make("U3", 1, parametrized=True)
CNOT = _make("CNOT", 2)
CY = _make("CY", 2)
CZ = _make("CZ", 2, diagonal=True)
CH = _make("CH", 2)
CSX = _make("CSX", 2)
CSXDG = _make("CSXDG", 2)
CRX = _make("CRX", 2, parametrized=True)
CRY = _make("CRY", 2, parametrized=True)
CRZ = _make("CRZ", 2, parametrized=True, diagonal=True)
CU1 = _make("CU1", 2, parametrized=True, diagonal=True)
CU2 = _make("CU2", 2, parametrized=True)
CU3 = _make("CU3", 2, parametrized=True)
TOFFOLI = _make("TOFFOLI", 3)
CCZ = _make("CCZ", 3, diagonal=True)
DEUTSCH = _make("DEUTSCH", 3, parametrized=True)
SWAP = _make("SWAP", 2, ntargets=2)
iSWAP = _make("iSWAP", 2, ntargets=2)
SiSWAP = _make("SiSWAP", 2, ntargets=2)
FSWAP = _make("FSWAP", 2, ntargets=2)
fSim = _make("fSim", 2, ntargets=2, parametrized=True)
RXX = _make("RXX", 2, ntargets=2, parametrized=True)
RYY = _make("RYY", 2, ntargets=2, parametrized=True)
RZZ = _make("RZZ", 2, ntargets=2, parametrized=True, diagonal=True)


class GeneralizedfSim(Gate):
    nqubit_args = 2
    ntargets = 2
    parametrized = True

    def __init__(self, q0, q1, unitary, phi):
        super().__init__(q0, q1)
        self.parameters = (np.asarray(unitary), phi)
        self.init_args = [q0, q1, unitary, phi]


class Unitary(Gate):
    """Arbitrary unitary (or any matrix, like the reference's tests use) on `targets`."""

    parametrized = True

    def __init__(self, unitary, *targets):
        self.ntargets = len(targets)
        self.nqubit_args = len(targets)
        super().__init__(*targets)
        u = np.asarray(unitary)
        dim = 1 << len(targets)
        if u.shape != (dim, dim):
            raise ValueError(f"Unitary on {len(targets)} qubits needs a {dim}x{dim} matrix")
        self.parameters = (u,)
        self.init_args = [unitary] + list(targets)
        self.name = "unitary"

    def target_matrix(self, matrices):
        return matrices.Unitary(self.parameters[0])


class FanOut(Gate):
    """One control, CNOT onto every other listed qubit (cpu.py:417-431)."""

    def __init__(self, control, *targets):
        self.ntargets = len(targets)
        self.nqubit_args = len(targets) + 1
        super().__init__(control, *targets)
        self.name = "fanout"


class FusedGate(Gate):
    """A block of gates acting on `targets`, applied as one dense matrix (qibo ``FusedGate``,
    consumed through ``matrix_fused`` at cpu.py:535-537)."""

    parametrized = False

    def __init__(self, *targets):
        self.ntargets = len(targets)
        self.nqubit_args = len(targets)
        super().__init__(*targets)
        self.gates = []
        self.name = "fused"
        self._matrix = None

    def append(self, gate):
        self.gates.append(gate)
        self._matrix = None


class M(Gate):
    """Measurement of `qubits` (collapse=True makes it a mid-circuit projective one)."""

    def __init__(self, *qubits, collapse=False):
        self.ntargets = len(qubits)
        self.nqubit_args = len(qubits)
        super().__init__(*qubits)
        self.collapse = collapse
        self.name = "measure"
        self.result = None

    def apply(self, backend, state, nqubits):
        if not self.collapse:
            return state
        qubits = sorted(self.target_qubits)
        probs = backend.calculate_probabilities(state, qubits, nqubits)
        shot = backend.sample_shots(probs, 1)
        self.result = int(backend.to_numpy(shot).ravel()[0])
        return backend.collapse_state(state, qubits, self.result, nqubits)
