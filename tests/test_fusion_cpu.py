"""CPU tests of the gate-fusion stand-in (qibo `Circuit.fuse` / `FusedGate` / `matrix_fused`,
consumed at /root/reference/src/qibojit/backends/cpu.py:535-537): a fused circuit must act like its
unfused gate list.  The checker is independent of the product's fused-matrix builder: the unfused
gates go one by one through the einsum oracle (tests/refdispatch.einsum_apply), the fused blocks
through `fusion.fused_matrix` + the same einsum."""

import numpy as np
import pytest

from qibojit_b200 import circuits, fusion, gates, planner
from qibojit_b200.circuit import Circuit
from qibojit_b200.matrices import CustomMatrices
from tests import refdispatch as R
from tests.circuits_random import random_circuit_gates

MATS = CustomMatrices("complex128")


def unfused_reference(state, queue, n):
    """Gate by gate with each gate's OWN matrix; FusedGate blocks are opened, never multiplied."""
    flat = []
    for g in queue:
        flat.extend(g.gates if g.__class__.__name__ == "FusedGate" else [g])
    assert not any(g.__class__.__name__ == "FusedGate" for g in flat)
    return R.reference_run(state, flat, n)


@pytest.mark.parametrize("max_qubits", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("seed", range(4))
def test_fused_circuit_equals_unfused(max_qubits, seed):
    n = 9
    glist = random_circuit_gates(n, 70, 100 + seed)
    c = Circuit(n)
    c.add(glist)
    fused = c.fuse(max_qubits=max_qubits)
    nblocks = sum(g.__class__.__name__ == "FusedGate" for g in fused.queue)
    assert max_qubits == 1 or (nblocks > 0 and len(fused.queue) < len(glist))
    st = R.random_state(n, "complex128", seed)
    ref = R.reference_run(st, glist, n)                               # the original gate list
    np.testing.assert_allclose(unfused_reference(st, fused.queue, n), ref, rtol=0, atol=1e-12)   # fusion kept the order
    got = np.array(st, dtype=np.complex128)
    for g in fused.queue:                                             # the fused matrices themselves
        m = fusion.target_only_matrix(g, MATS)
        got = R.einsum_apply(got, m, list(g.target_qubits), list(g.control_qubits), n)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["qft", "variational", "supremacy", "quantum_volume"])
def test_fused_benchmark_circuits_equal_unfused(name):
    n = 10
    kw = {"depth": 3} if name in ("supremacy", "quantum_volume") else {}
    c = getattr(circuits, name)(n, **kw)
    st = R.random_state(n, "complex128", 3)
    ref = R.reference_run(st, c.queue, n)
    for k in (2, 4):
        fused = c.fuse(max_qubits=k)
        got = np.array(st, dtype=np.complex128)
        for g in fused.queue:
            got = R.einsum_apply(got, fusion.target_only_matrix(g, MATS), list(g.target_qubits),
                                 list(g.control_qubits), n)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


def test_pair_blocks_of_the_planner_equal_the_separate_ops():
    """planner.fuse_pair_blocks (RY RY CZ RY RY -> one real 4x4): the op list with blocks acts like
    the op list without, checked with op matrices on random states."""
    n = 8
    for seed, glist in enumerate([circuits.variational(n).queue, circuits.supremacy(n, depth=4).queue,
                                  random_circuit_gates(n, 80, 5), circuits.qft(n).queue]):
        ops = []
        for g in glist:
            ops.extend(planner.lower_gate(g, n, MATS))
        had_raw = any(op.kind == "raw" for op in ops)      # (dense gates on >= 3 qubits are not planner ops)
        ops = [op for op in ops if op.kind != "raw"]
        fused = planner.fuse_pair_blocks(planner.fuse_one_qubit_runs(ops))

        def run(oplist, st):
            st = np.array(st, dtype=np.complex128)
            for op in oplist:
                bits = sorted(op.bits)
                m = planner.op_matrix(op, bits)
                # index bit b <-> qubit n-1-b; einsum wants the most significant matrix bit first
                st = R.einsum_apply(st, m, [n - 1 - b for b in reversed(bits)], [], n)
            return st

        st = R.random_state(n, "complex128", seed)
        np.testing.assert_allclose(run(fused, st), run(ops, st), rtol=0, atol=1e-12)
        if not had_raw:
            np.testing.assert_allclose(run(ops, st), R.reference_run(st, glist, n), rtol=0, atol=1e-12)
    var = []
    for g in circuits.variational(n).queue:
        var.extend(planner.lower_gate(g, n, MATS))
    blocks = [op for op in planner.fuse_pair_blocks(var) if op.kind == "dense" and len(op.targets) == 2]
    assert len(blocks) == 2 * (n // 2) and all(not np.any(np.asarray(op.data).imag) for op in blocks)


def test_fused_matrix_cache_follows_dtype_and_parameters():
    fg = gates.FusedGate(0, 1)
    rx = gates.RX(0, 0.3)
    fg.append(rx)
    fg.append(gates.CZ(0, 1))
    m64 = fusion.fused_matrix(fg, CustomMatrices("complex64"))
    m128 = fusion.fused_matrix(fg, CustomMatrices("complex128"))
    assert m64.dtype == np.complex64 and m128.dtype == np.complex128
    exact = fusion.full_matrix(gates.CZ(0, 1), MATS) @ np.kron(fusion.target_only_matrix(rx, MATS), np.eye(2))
    np.testing.assert_allclose(m128, exact, rtol=0, atol=1e-15)        # not a widened float32 matrix
    rx.parameters = (0.9,)                                             # re-parametrised inner gate: rebuilt
    exact2 = fusion.full_matrix(gates.CZ(0, 1), MATS) @ np.kron(fusion.target_only_matrix(rx, MATS), np.eye(2))
    np.testing.assert_allclose(fusion.fused_matrix(fg, MATS), exact2, rtol=0, atol=1e-15)
    assert np.abs(exact2 - exact).max() > 0.1
