"""Shared helpers of the -m gpu parity tests (CUDA path through the C ABI vs the oracle)."""

import functools

import pytest

from oracle import oracle as O
from tests import refdispatch as R

# parity bar of BASELINE.json's north star
ATOL = {"complex64": 1e-5, "complex128": 1e-12}

ORACLE_DISPATCH = (
    functools.partial(R.one_qubit_base, O),
    functools.partial(R.two_qubit_base, O),
    functools.partial(R.multi_qubit_base, O),
)


@functools.lru_cache(maxsize=1)
def backend():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from qibojit_b200.backends.b200 import B200Backend

    return B200Backend()


def gpu_dispatch(b):
    """(one, two, multi) callables with the reference's `_x_qubit_base` signatures that take a
    numpy state, run the CUDA kernels and hand back numpy."""

    def one(st, nq, target, kernel, gate, qubits):
        d = b.cast(st, dtype=str(st.dtype), copy=True)
        return b.to_numpy(b._one_qubit_base(d, nq, target, kernel, gate, qubits))

    def two(st, nq, t1, t2, kernel, gate, qubits):
        d = b.cast(st, dtype=str(st.dtype), copy=True)
        return b.to_numpy(b._two_qubit_base(d, nq, t1, t2, kernel, gate, qubits))

    def multi(st, nq, targets, gate, qubits):
        d = b.cast(st, dtype=str(st.dtype), copy=True)
        return b.to_numpy(b._multi_qubit_base(d, nq, targets, gate, qubits))

    return one, two, multi
