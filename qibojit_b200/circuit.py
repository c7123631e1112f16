"""Tiny circuit container with the attributes the backend reads from a qibo ``Circuit``
(``nqubits``, ``queue``)."""

from . import fusion


class Circuit:
    def __init__(self, nqubits):
        self.nqubits = int(nqubits)
        self.queue = []

    def add(self, gate):
        if isinstance(gate, (list, tuple)) or hasattr(gate, "__next__"):
            for g in gate:
                self.add(g)
            return
        for q in gate.qubits:
            if not 0 <= q < self.nqubits:
                raise ValueError(f"gate {gate} acts outside a {self.nqubits}-qubit circuit")
        self.queue.append(gate)

    @property
    def ngates(self):
        return len(self.queue)

    def fuse(self, max_qubits=2):
        out = Circuit(self.nqubits)
        out.queue = fusion.fuse(self.queue, max_qubits=max_qubits)
        return out

    def __call__(self, backend, initial_state=None):
        return backend.execute_circuit(self, initial_state=initial_state)
