"""-m gpu parity tests of the gate kernels: CUDA path (through the C ABI) vs the committed
golden vectors of the reference's numba kernels and vs the oracle, on the reference's own
parametrisation tables (tests/cases.py).  Bar: 1e-12 (complex128) / 1e-5 (complex64) max-abs."""

import numpy as np
import pytest

from tests import cases, goldenio
from tests import refdispatch as R
from tests.gpu_utils import ATOL, ORACLE_DISPATCH, backend, gpu_dispatch

pytestmark = pytest.mark.gpu

GATE_CASES = list(goldenio.iter_gate_cases(cases, R))


@pytest.mark.parametrize("route", [1, 0, 2], ids=["direct", "auto", "tile"])
@pytest.mark.parametrize("case", GATE_CASES, ids=[c[0] for c in GATE_CASES])
def test_gate_kernels_match_reference_golden(case, route, golden_gates):
    key, kind, dtype, nq, t, c, seed = case
    b = backend()
    b.set_route(route)
    try:
        out = goldenio.run_gate_case(gpu_dispatch(b), R, kind, dtype, nq, t, c, seed)
    finally:
        b.set_route(0)
    if kind == "multilarge":
        np.testing.assert_allclose(goldenio.subsample(out, seed), golden_gates[key],
                                   atol=1e-3 if dtype == "complex64" else 1e-10)
        ref = goldenio.run_gate_case(ORACLE_DISPATCH, R, kind, dtype, nq, t, c, seed)
        np.testing.assert_allclose(out, ref, rtol=0, atol=ATOL[dtype])
    else:
        np.testing.assert_allclose(out, golden_gates[key], rtol=0, atol=ATOL[dtype])


@pytest.mark.parametrize("dtype", cases.DTYPES)
@pytest.mark.parametrize("use_qubits", [False, True])
@pytest.mark.parametrize(("nqubits", "target"), [(4, 1), (6, 5)])
def test_one_qubit_base(nqubits, target, use_qubits, dtype):
    # /root/reference/tests/test_gates.py:52-68
    b = backend()
    st = R.random_state(nqubits, dtype, 3)
    m = R.random_matrix(2, dtype, 3)
    expect = R.einsum_apply(st, m, [target], [], nqubits)
    qubits = R.qubits_tensor(nqubits, [target]) if use_qubits else None
    out = b._one_qubit_base(b.cast(st, dtype=dtype), nqubits, target, "apply_gate", m, qubits)
    b.assert_allclose(out, expect, atol=1e-4 if dtype == "complex64" else 1e-10)


@pytest.mark.parametrize("dtype", cases.DTYPES)
@pytest.mark.parametrize("use_qubits", [False, True])
@pytest.mark.parametrize(("nqubits", "targets"), [(5, [3, 4]), (4, [2, 0])])
def test_two_qubit_base(nqubits, targets, use_qubits, dtype):
    # /root/reference/tests/test_gates.py:144-160
    b = backend()
    st = R.random_state(nqubits, dtype, 4)
    m = R.random_matrix(4, dtype, 4)
    expect = R.einsum_apply(st, m, targets, [], nqubits)
    qubits = R.qubits_tensor(nqubits, targets) if use_qubits else None
    out = b._two_qubit_base(b.cast(st, dtype=dtype), nqubits, *targets, "apply_two_qubit_gate", m, qubits)
    b.assert_allclose(out, expect, atol=1e-4 if dtype == "complex64" else 1e-10)


@pytest.mark.parametrize("dtype", cases.DTYPES)
@pytest.mark.parametrize("use_qubits", [False, True])
@pytest.mark.parametrize(("nqubits", "targets"), [(5, [2, 3, 4]), (4, [2, 0, 1])])
def test_multi_qubit_base(nqubits, targets, use_qubits, dtype):
    # /root/reference/tests/test_gates.py:349-366
    b = backend()
    st = R.random_state(nqubits, dtype, 5)
    m = R.random_matrix(8, dtype, 5)
    expect = R.einsum_apply(st, m, targets, [], nqubits)
    qubits = R.qubits_tensor(nqubits, targets) if use_qubits else None
    out = b._multi_qubit_base(b.cast(st, dtype=dtype), nqubits, targets, m, qubits)
    b.assert_allclose(out, expect, atol=1e-4 if dtype == "complex64" else 1e-10)


def test_too_many_targets_raises():
    # gpu.py:989-993
    b = backend()
    n = 12
    st = b.zero_state(n)
    with pytest.raises(ValueError):
        b._multi_qubit_base(st, n, list(range(11)), np.eye(2 ** 11), None)


@pytest.mark.parametrize("route", [1, 2], ids=["direct", "tile"])
@pytest.mark.parametrize("dtype", cases.DTYPES)
@pytest.mark.parametrize("nqubits", [18, 22])
def test_every_target_bit_vs_oracle(nqubits, dtype, route):
    """Sweep the target over every index bit (all access regimes of the kernels)."""
    b = backend()
    from oracle import oracle as O

    st = R.random_state(nqubits, dtype, 9)
    d = b.cast(st, dtype=dtype, copy=True)
    ref = st.copy()
    b.set_route(route)
    try:
        for target in range(nqubits):
            m = R.random_matrix(2, dtype, target)
            m = m / np.linalg.norm(m, 2)
            d = b._one_qubit_base(d, nqubits, target, "apply_gate", m, None)
            ref = R.one_qubit_base(O, ref, nqubits, target, "apply_gate", m, None)
    finally:
        b.set_route(0)
    b.assert_allclose(d, ref, rtol=0, atol=ATOL[dtype])


@pytest.mark.parametrize("route", [1, 2], ids=["direct", "tile"])
@pytest.mark.parametrize("dtype", cases.DTYPES)
@pytest.mark.parametrize("k", [2, 3, 4, 5])
def test_random_target_sets_vs_oracle(k, dtype, route):
    """Dense k-target gates on random (unsorted) target sets with random controls, n = 20:
    low, mixed and high bit positions; the fused-block kernel's whole geometry space."""
    b = backend()
    from oracle import oracle as O

    n = 20
    rng = np.random.default_rng(100 + k)
    st = R.random_state(n, dtype, 13)
    d = b.cast(st, dtype=dtype, copy=True)
    ref = st.copy()
    b.set_route(route)
    try:
        for trial in range(8):
            qs = [int(v) for v in rng.permutation(n)[: k + (trial % 3)]]
            targets, controls = qs[:k], qs[k:]
            if trial == 0:
                targets = list(range(n - k, n))          # the k lowest index bits
                controls = []
            if trial == 1:
                targets = list(range(k))[::-1]            # the k highest index bits, reversed
                controls = [n - 1]
            m = R.random_matrix(1 << k, dtype, trial)
            m = m / np.linalg.norm(m, 2)
            q = R.qubits_tensor(n, targets, controls)
            d = b._multi_qubit_base(d, n, targets, m, q)
            ref = R.multi_qubit_base(O, ref, n, targets, m, q)
    finally:
        b.set_route(0)
    b.assert_allclose(d, ref, rtol=0, atol=ATOL[dtype])


@pytest.mark.parametrize("dtype", cases.DTYPES)
def test_gate_objects_dispatch(dtype):
    """GATE_OPS dispatch through `apply_gate` on gate objects (cpu.py:23-36, 433-450)."""
    from qibojit_b200 import gates
    from qibojit_b200.matrices import CustomMatrices
    from qibojit_b200 import fusion

    b = backend()
    b.set_dtype(dtype)
    n = 7
    st = R.random_state(n, dtype, 21)
    glist = [
        gates.H(0), gates.X(3), gates.Y(6), gates.Z(2), gates.CNOT(1, 4), gates.CZ(6, 0),
        gates.CY(2, 5), gates.TOFFOLI(0, 6, 3), gates.U1(4, 0.3), gates.CU1(5, 1, 0.7),
        gates.SWAP(0, 6), gates.SWAP(3, 2).controlled_by(5), gates.fSim(1, 5, 0.4, 0.9),
        gates.GeneralizedfSim(6, 2, R.random_matrix(2, dtype, 1), 0.33),
        gates.RX(1, 0.2), gates.RY(6, 1.2).controlled_by(0, 3), gates.CRZ(2, 4, 0.5),
        gates.CH(5, 6), gates.CSX(1, 0), gates.CCZ(1, 2, 6), gates.DEUTSCH(0, 1, 2, 0.4),
        gates.iSWAP(4, 1), gates.RZZ(0, 6, 0.8), gates.Unitary(R.random_matrix(8, dtype, 2), 5, 0, 3),
        gates.FanOut(2, 0, 5, 6),
    ]
    d = b.cast(st, dtype=dtype, copy=True)
    ref = st.astype(np.complex128)
    mats = CustomMatrices("complex128")
    for g in glist:
        d = b.apply_gate(g, d, n)
        if g.name == "fanout":
            for tq in g.target_qubits:
                ref = R.einsum_apply(ref, mats.X, [tq], [g.control_qubits[0]], n)
        else:
            ref = R.einsum_apply(ref, fusion.target_only_matrix(g, mats), list(g.target_qubits),
                                 list(g.control_qubits), n)
    b.assert_allclose(d, ref, rtol=0, atol=2e-4 if dtype == "complex64" else 1e-10)
    b.set_dtype("complex128")
