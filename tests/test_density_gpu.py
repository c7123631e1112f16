"""-m gpu parity tests of the density-matrix path: the doubling trick of
/root/reference/src/qibojit/backends/cpu.py:452-517 on the CUDA kernels, `apply_gate_half_density_matrix`
and channels, against U rho U^dagger computed with full matrices in numpy -- the role qibo's
NumpyBackend plays in the reference's own tests (tests/test_gates.py:112-141, 295-329, 414-453)."""

import types

import numpy as np
import pytest

from qibojit_b200 import fusion, gates
from qibojit_b200.matrices import CustomMatrices
from tests import refdispatch as R
from tests.gpu_utils import backend

pytestmark = pytest.mark.gpu

ATOL = {"complex128": 1e-10, "complex64": 1e-4}      # the reference's own bar (tests/test_gates.py:15)
MATS = CustomMatrices("complex128")


def random_density_matrix(n, dtype, seed):
    rng = np.random.default_rng(seed)
    d = 1 << n
    a = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
    rho = a @ a.conj().T
    return (rho / np.trace(rho)).astype(dtype)


def full_unitary(gate, n):
    """2^n x 2^n matrix of `gate` (controls included), via the einsum oracle on basis vectors."""
    d = 1 << n
    u = np.zeros((d, d), dtype=np.complex128)
    for j in range(d):
        e = np.zeros(d, dtype=np.complex128)
        e[j] = 1
        u[:, j] = R.reference_run(e, [gate], n)
    return u


def run_dm(b, gate, rho, n, dtype, **kw):
    b.set_dtype(dtype)
    try:
        d = b.cast(rho, dtype=dtype, copy=True)
        out = b.apply_gate(gate, d, n, **kw)
        return b.to_numpy(out)
    finally:
        b.set_dtype("complex128")


def gate_cases():
    u2 = R.random_unitary(2, 1)
    u4 = R.random_unitary(4, 2)
    u8 = R.random_unitary(8, 3)
    u16 = R.random_unitary(16, 4)
    return [
        (3, gates.H(1)), (3, gates.X(0)), (3, gates.Y(2)), (3, gates.Z(1)), (4, gates.U1(2, 0.1234)),
        (4, gates.RX(3, 0.3)), (4, gates.Unitary(u2, 1)), (5, gates.Unitary(u2, 3).controlled_by(0, 4)),
        (4, gates.CNOT(0, 2)), (4, gates.CY(3, 1)), (4, gates.CZ(1, 2)), (4, gates.CU1(0, 3, 0.7)),
        (4, gates.TOFFOLI(0, 1, 3)), (4, gates.X(2).controlled_by(0, 1, 3)),
        (5, gates.Unitary(u4, 3, 4)), (4, gates.Unitary(u4, 2, 0)), (5, gates.Unitary(u4, 3, 1).controlled_by(0, 2)),
        (4, gates.SWAP(0, 3)), (4, gates.SWAP(1, 2).controlled_by(0)), (4, gates.fSim(1, 3, 0.4, 0.9)),
        (4, gates.GeneralizedfSim(2, 0, R.random_unitary(2, 9), 0.6)), (4, gates.RZZ(0, 2, 0.3)),
        (4, gates.Unitary(u8, 2, 1, 3)), (5, gates.Unitary(u8, 0, 2, 3).controlled_by(1)),
        (5, gates.Unitary(u16, 0, 2, 3, 4)), (4, gates.FanOut(0, 1, 3)),
    ]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("case", range(len(gate_cases())))
def test_density_matrix_gate(case, dtype):
    """U rho U^dagger through `apply_gate` on a 2-D state (cpu.py:369-377 -> 452-517)."""
    n, gate = gate_cases()[case]
    b = backend()
    rho = random_density_matrix(n, dtype, case)
    u = full_unitary(gate, n)
    ref = u @ rho.astype(np.complex128) @ u.conj().T
    got = run_dm(b, gate, rho, n, dtype)
    assert got.shape == rho.shape
    np.testing.assert_allclose(got, ref, rtol=0, atol=ATOL[dtype])


@pytest.mark.parametrize("case", range(len(gate_cases())))
def test_density_matrix_inverse_gate(case):
    """`inverse=True` (the reset step of unitary channels, cpu.py:464-468): U^-1 rho U^-1^dagger,
    including the gates whose kernel buffer is not a matrix (U1: scalar, fSim: 5-vector)."""
    n, gate = gate_cases()[case]
    b = backend()
    rho = random_density_matrix(n, "complex128", 50 + case)
    u = np.linalg.inv(full_unitary(gate, n))
    ref = u @ rho @ u.conj().T
    got = run_dm(b, gate, rho, n, "complex128", inverse=True)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)


@pytest.mark.parametrize("gatename", ["H", "X", "Z", "Y"])
def test_density_matrix_half_calls(gatename):
    """tests/test_gates.py:414-422: only U acting on the row index, U rho."""
    b = backend()
    rho = random_density_matrix(3, "complex128", 7)
    gate = getattr(gates, gatename)(1)
    ref = full_unitary(gate, 3) @ rho
    got = b.to_numpy(b.apply_gate_half_density_matrix(gate, b.cast(rho, copy=True), 3))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)
    # a controlled two-qubit block as well
    gate = gates.Unitary(R.random_unitary(4, 5), 2, 0).controlled_by(1)
    got = b.to_numpy(b.apply_gate_half_density_matrix(gate, b.cast(rho, copy=True), 3))
    np.testing.assert_allclose(got, full_unitary(gate, 3) @ rho, rtol=0, atol=1e-10)


def _channel(pairs, all_unitary):
    ch = types.SimpleNamespace()
    ch.coefficients = tuple(p for p, _ in pairs)
    ch.gates = tuple(g for _, g in pairs)
    ch.coefficient_sum = float(sum(ch.coefficients))
    ch._all_unitary_operators = all_unitary
    return ch


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_unitary_channel(dtype):
    """tests/test_gates.py:425-441: (1 - sum p) rho + sum p_k U_k rho U_k^dagger, state restored by
    the inverse gate between the terms."""
    b = backend()
    n = 4
    a1 = gates.Unitary(np.asarray(MATS.X), 0)
    a2 = gates.Unitary(np.asarray(fusion.full_matrix(gates.CNOT(0, 1), MATS)), 2, 3)
    pairs = [(0.4, a1), (0.3, a2)]
    rho = random_density_matrix(n, dtype, 10)
    ref = (1 - 0.7) * rho.astype(np.complex128)
    for p, g in pairs:
        u = full_unitary(g, n)
        ref = ref + p * (u @ rho.astype(np.complex128) @ u.conj().T)
    b.set_dtype(dtype)
    try:
        got = b.to_numpy(b.apply_channel(_channel(pairs, True), b.cast(rho, dtype=dtype, copy=True), n))
    finally:
        b.set_dtype("complex128")
    np.testing.assert_allclose(got, ref, rtol=0, atol=ATOL[dtype])


def test_unitary_channel_with_kernel_format_gates():
    """Gates whose kernel buffer is a scalar / 5-vector / absent (U1, fSim, FanOut, Y) inside a
    unitary channel: the inverse must come from the gate's matrix."""
    b = backend()
    n = 4
    pairs = [(0.2, gates.U1(1, 0.3)), (0.1, gates.fSim(0, 2, 0.5, 0.2)), (0.15, gates.CU1(3, 0, 1.1)),
             (0.1, gates.Y(2)), (0.05, gates.FanOut(1, 0, 3)), (0.1, gates.SWAP(1, 3))]
    rho = random_density_matrix(n, "complex128", 11)
    ref = (1 - sum(p for p, _ in pairs)) * rho
    for p, g in pairs:
        u = full_unitary(g, n)
        ref = ref + p * (u @ rho @ u.conj().T)
    got = b.to_numpy(b.apply_channel(_channel(pairs, True), b.cast(rho, copy=True), n))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)


def test_non_unitary_channel_restores_a_copy():
    """cpu.py:403-413: without `_all_unitary_operators` the state is restored from a copy."""
    b = backend()
    n = 3
    k0 = np.array([[1, 0], [0, np.sqrt(0.7)]], dtype=np.complex128)
    k1 = np.array([[0, np.sqrt(0.3)], [0, 0]], dtype=np.complex128)
    pairs = [(1.0, gates.Unitary(k0, 1)), (1.0, gates.Unitary(k1, 1))]
    rho = random_density_matrix(n, "complex128", 12)
    ch = _channel(pairs, False)
    ch.coefficient_sum = 1.0          # Kraus channel: (1 - 1) rho + sum K rho K^dagger
    ref = np.zeros_like(rho)
    for _, g in pairs:
        u = full_unitary(g, n)
        ref = ref + u @ rho @ u.conj().T
    got = b.to_numpy(b.apply_channel(ch, b.cast(rho, copy=True), n))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)
    assert abs(np.trace(got) - 1) < 1e-10


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("qubits,shot", [([0], 1), ([1, 2], 2), ([0, 3], 3), ([0, 1, 2, 3], 9), ([2], 0)])
def test_density_matrix_collapse_and_probabilities(qubits, shot, dtype):
    """qibo `collapse_density_matrix` / `calculate_probabilities_density_matrix` restated with numpy
    projectors: P rho P / tr(P rho), and the marginal of the diagonal."""
    b = backend()
    n = 4
    rho = random_density_matrix(n, dtype, 13).astype(np.complex128)
    d = 1 << n
    idx = np.arange(d)
    sel = np.ones(d, dtype=bool)
    for j, q in enumerate(qubits):
        sel &= ((idx >> (n - 1 - q)) & 1) == ((shot >> (len(qubits) - 1 - j)) & 1)
    proj = np.diag(sel.astype(np.complex128))
    ref = proj @ rho @ proj
    ref_un = ref.copy()
    ref = ref / np.trace(ref)
    b.set_dtype(dtype)
    try:
        got = b.to_numpy(b.collapse_state(b.cast(rho, dtype=dtype, copy=True), qubits, shot, n, density_matrix=True))
        got_un = b.to_numpy(b.collapse_state(b.cast(rho, dtype=dtype, copy=True), qubits, shot, n, normalize=False,
                                             density_matrix=True))
        probs = b.to_numpy(b.calculate_probabilities(b.cast(rho, dtype=dtype, copy=True), qubits, n,
                                                     density_matrix=True))
    finally:
        b.set_dtype("complex128")
    np.testing.assert_allclose(got, ref, rtol=0, atol=ATOL[dtype])
    np.testing.assert_allclose(got_un, ref_un, rtol=0, atol=ATOL[dtype])
    diag = np.real(np.diagonal(rho)).reshape((2,) * n)
    rest = tuple(q for q in range(n) if q not in qubits)
    marg = diag.sum(axis=rest) if rest else diag
    kept = sorted(qubits)
    marg = np.transpose(marg, [kept.index(q) for q in qubits]).reshape(-1)
    np.testing.assert_allclose(probs, np.abs(marg), rtol=0, atol=ATOL[dtype])


def test_zero_density_matrix_then_circuit():
    """initial_density_matrix (ops.py:25-30) + a few gates = |psi><psi| of the state-vector run."""
    b = backend()
    n = 3
    glist = [gates.H(0), gates.CNOT(0, 1), gates.RY(2, 0.4), gates.CU1(2, 0, 0.3), gates.SWAP(1, 2)]
    rho = b.zero_state(n, density_matrix=True)
    psi = b.zero_state(n)
    for g in glist:
        rho = b.apply_gate(g, rho, n)
        psi = b.apply_gate(g, psi, n)
    v = b.to_numpy(psi)
    np.testing.assert_allclose(b.to_numpy(rho), np.outer(v, v.conj()), rtol=0, atol=1e-12)
