"""CPU tests of the pass planner (host logic only): partition + round schedule + diagonal merging
must be a re-ordering of memory traffic, not of the circuit's meaning."""

import numpy as np
import pytest

from qibojit_b200 import circuits, gates, planner
from qibojit_b200.matrices import CustomMatrices
from tests import plan_interp
from tests import refdispatch as R
from tests.circuits_random import random_circuit_gates

MATS = CustomMatrices("complex128")


def _raw(n):
    def apply_raw(state, gate):
        return R.reference_run(state, [gate], n)
    return apply_raw


def _ops(plan):
    return [op for seg in plan if seg[0] == "pass" for _, ops in seg[2] for op in ops]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("n,tile_bits,run_bits", [(6, 6, 2), (8, 6, 3), (9, 7, 2), (10, 10, 5), (7, 12, 5)])
def test_plan_matches_gate_by_gate(n, tile_bits, run_bits, seed, dtype):
    glist = random_circuit_gates(n, 60, seed)
    st = R.random_state(n, "complex128", seed)
    plan = planner.plan_queue(glist, n, MATS, tile_bits, run_bits, max_diag_bits=4 + seed % 3, dtype=dtype)
    got = plan_interp.run_plan(st.copy(), plan, n, _raw(n), nreg=planner.REG_BITS[dtype],
                               fixed=(0,) if dtype == "complex64" else ())
    ref = R.reference_run(st, glist, n)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


def test_every_dense_target_is_local_and_passes_are_well_formed():
    n = 12
    glist = random_circuit_gates(n, 200, 3)
    plan = planner.plan_queue(glist, n, MATS, 7, 3)
    for seg in plan:
        if seg[0] != "pass":
            continue
        local = seg[1]
        assert len(local) == 7 and local == sorted(set(local)) and local[:3] == [0, 1, 2]
        for regs, ops in seg[2]:
            assert len(regs) == 4 and set(regs) <= set(local)
            for op in ops:
                assert len(set(op.targets) | set(op.controls)) == len(op.targets) + len(op.controls)
                if op.kind == "dense":
                    assert set(op.targets) <= set(regs) and len(op.targets) in (1, 2)
                else:
                    assert len(op.targets) <= 10


def test_qft_needs_few_passes_and_merges_its_phase_ladders():
    n = 33
    plan = planner.plan_queue(circuits.qft(n).queue, n, MATS, 12, 5)
    assert all(seg[0] == "pass" for seg in plan)
    assert len(plan) <= 9                      # 577 gates, 577 passes gate by gate
    assert len(_ops(plan)) < 577 // 2          # 528 CU1 merged into far fewer tables
    # a phase ladder's table touches at most two register bits of its round
    for seg in plan:
        for regs, ops in seg[2]:
            for op in ops:
                if op.kind == "diag":
                    assert len(op.bits & set(regs)) <= 2


def test_qft_plan_small_matches_analytic():
    n = 10
    plan = planner.plan_queue(circuits.qft(n).queue, n, MATS, 6, 3)
    st = np.zeros(1 << n, dtype=np.complex128)
    st[0] = 1
    got = plan_interp.run_plan(st, plan, n, _raw(n))
    np.testing.assert_allclose(got, np.full(1 << n, 2.0 ** (-n / 2)), rtol=0, atol=1e-14)


def test_diagonal_controls_are_split_off():
    n = 10   # all four qubits lie outside the 6-bit tile: one table family
    glist = [gates.CU1(1, 0, 0.3), gates.CU1(2, 0, 0.5), gates.CU1(3, 0, 0.7)]
    plan = planner.plan_queue(glist, n, MATS, 6, 3)
    assert len(plan) == 1 and len(_ops(plan)) == 1
    op = _ops(plan)[0]
    assert op.kind == "diag" and op.controls == (n - 1,) and len(op.targets) == 3


def test_rounds_hoist_commuting_gates():
    """Gates on disjoint qubits are packed into one round up to the register budget."""
    n = 12
    glist = [gates.H(q) for q in range(n)]
    plan = planner.plan_queue(glist, n, MATS, 12, 5)
    assert len(plan) == 1 and len(plan[0][2]) == 3      # 12 H gates, 4 register bits per round
    plan64 = planner.plan_queue(glist, n, MATS, 12, 5, dtype="complex64")
    assert len(plan64[0][2]) == 3 and all(0 in regs for regs, _ in plan64[0][2])


def test_phase_tables_are_split_by_locality():
    """Within a round, tables over tile-local bits, over outer bits and over both are kept apart
    (the kernel turns the first two kinds into one factor per thread / per tile)."""
    n = 16
    plan = planner.plan_queue(circuits.qft(n).queue, n, MATS, 8, 3)
    for seg in plan:
        local = set(seg[1])
        for regs, ops in seg[2]:
            for op in ops:
                if op.kind != "diag":
                    continue
                rest = op.bits - set(regs)
                assert rest <= local or not (rest & local), (sorted(rest), sorted(local))


def test_fma_count_of_the_benchmark_circuits():
    """The arithmetic side of the roofline bench.py reports: 4 real FMAs per amplitude for a real
    or axis-aligned one-qubit gate, 8 for a complex one, 16 for a 4x4, none for sign flips."""
    class Stub(planner.Program):
        def __init__(self, queue, n, dt, pair_blocks=True):
            self.passes = [(s[1], s[2]) for s in planner.plan_queue(
                queue, n, CustomMatrices(dt), planner.DEFAULT_TILE_BITS[dt], planner.DEFAULT_RUN_BITS[dt], 10, dt,
                pair_blocks=pair_blocks)
                if s[0] == "pass"]
            self.segments = []

    n = 20
    var = circuits.variational(n).queue
    assert Stub(var, n, "complex128", pair_blocks=False).fma_per_amplitude() == 5 * n * 4    # RY real, CZ = signs
    # RY RY CZ RY RY on a pair = one real 4x4 (8 per amplitude instead of 16): two block layers + the last RY layer
    assert Stub(var, n, "complex128").fma_per_amplitude() == 2 * (n // 2) * 8 + n * 4
    qv = circuits.quantum_volume(n, depth=3)
    assert Stub(qv.queue, n, "complex64").fma_per_amplitude() == len(qv.queue) * 16
    assert Stub([gates.X(0), gates.CNOT(1, 2), gates.SWAP(3, 4), gates.Z(5)], n, "complex128").fma_per_amplitude() == 0


def test_local_gates_of_the_distributed_layer_lower_like_their_originals():
    """distributed.LocalGate carries a dense target matrix: the planner must classify it
    (diagonal / dense / raw) exactly as it does the gate it came from."""
    from qibojit_b200 import fusion
    from qibojit_b200.distributed import LocalGate

    n = 10
    for g in [gates.H(2), gates.CU1(1, 3, 0.3), gates.RZ(4, 0.2), gates.CNOT(0, 5), gates.fSim(1, 2, 0.3, 0.4),
              gates.TOFFOLI(0, 1, 2), gates.RZZ(3, 6, 0.5)]:
        dense = fusion.target_only_matrix(g, MATS)
        lg = LocalGate(None, g.target_qubits, g.control_qubits, dense, dense)
        a = planner.lower_gate(g, n, MATS)
        b = planner.lower_gate(lg, n, MATS)
        assert [op.kind for op in a] == [op.kind for op in b]
        for x, y in zip(a, b):
            assert x.targets == y.targets and x.controls == y.controls
            np.testing.assert_allclose(np.asarray(x.data), np.asarray(y.data), atol=1e-15)


@pytest.mark.parametrize("seed", range(6))
def test_swaps_are_relabelled_away_from_the_zero_state(seed):
    """zero_state plans contain no SWAP and give the same final state from |0...0>."""
    n = 9
    rng = np.random.default_rng(seed)
    glist = []
    for g in random_circuit_gates(n, 50, seed + 100):
        if len(g.target_qubits) > 2:           # raw gates switch the optimisation off (tested below)
            continue
        glist.append(g)
        if rng.random() < 0.25:
            a, b = rng.choice(n, size=2, replace=False)
            glist.append(gates.SWAP(int(a), int(b)))
    glist.append(gates.SWAP(0, n - 1))
    plan = planner.plan_queue(glist, n, MATS, 7, 3, zero_state=True)
    for op in _ops(plan):
        assert not (op.kind == "dense" and len(op.targets) == 2 and not op.controls
                    and np.array_equal(np.asarray(op.data).reshape(4, 4), planner._SWAP_MATRIX))
    st = np.zeros(1 << n, dtype=np.complex128)
    st[0] = 1
    got = plan_interp.run_plan(st.copy(), plan, n, _raw(n))
    ref = R.reference_run(st, glist, n)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


def test_qft_from_zero_state_needs_no_swap_passes():
    n = 33
    q = circuits.qft(n).queue
    with_swaps = planner.plan_queue(q, n, MATS, 11, 5)
    without = planner.plan_queue(q, n, MATS, 11, 5, zero_state=True)
    assert len(without) <= len(with_swaps) - 3
    small = circuits.qft(10).queue
    st = np.zeros(1 << 10, dtype=np.complex128)
    st[0] = 1
    got = plan_interp.run_plan(st, planner.plan_queue(small, 10, MATS, 6, 3, zero_state=True), 10, _raw(10))
    np.testing.assert_allclose(got, np.full(1 << 10, 2.0 ** -5), rtol=0, atol=1e-14)


def test_raw_gates_keep_the_swaps():
    n = 8
    u = np.linalg.qr(np.random.default_rng(1).standard_normal((8, 8)))[0]
    glist = [gates.H(0), gates.Unitary(u, 0, 3, 5), gates.SWAP(1, 2)]
    plan = planner.plan_queue(glist, n, MATS, 6, 3, zero_state=True)
    assert any(op.kind == "dense" and len(op.targets) == 2 for op in _ops(plan))


def test_adjacent_one_qubit_gates_are_fused():
    n = 8
    glist = [gates.H(0), gates.RX(0, 0.3), gates.RZ(0, 0.2), gates.CZ(0, 1), gates.H(0), gates.H(1), gates.T(1),
             gates.RY(1, 0.4), gates.Z(2), gates.S(2)]
    plan = planner.plan_queue(glist, n, MATS, 6, 3, pair_blocks=False)
    ops = _ops(plan)
    dense = [op for op in ops if op.kind == "dense"]
    assert len(dense) == 3                     # H.RX.RZ on qubit 0 | H on 0 after the CZ | H.T.RY on 1
    st = R.random_state(n, "complex128", 2)
    got = plan_interp.run_plan(st.copy(), plan, n, _raw(n))
    np.testing.assert_allclose(got, R.reference_run(st, glist, n), rtol=0, atol=1e-13)
    # phases alone stay phases (they merge into tables)
    assert all(op.kind == "diag" for op in _ops(planner.plan_queue([gates.Z(2), gates.S(2), gates.T(2)], n, MATS, 6, 3)))
    sup = circuits.supremacy(12, depth=4)
    assert len(_ops(planner.plan_queue(sup.queue, 12, MATS, 8, 3))) <= len(sup.queue) - 12   # H + first-cycle gate


def test_program_cache_key_follows_gate_parameters():
    """B200Backend.compile_circuit caches the program on the circuit object: the key must change
    when a gate parameter (or a Unitary's matrix) changes, because matrices are baked into the image."""
    from qibojit_b200.backends.b200 import B200Backend

    fp = B200Backend.circuit_fingerprint
    a = [gates.RY(0, 0.3), gates.CZ(0, 1), gates.Unitary(np.eye(2), 2)]
    b = [gates.RY(0, 0.3), gates.CZ(0, 1), gates.Unitary(np.eye(2), 2)]
    assert fp(a) == fp(b)
    b[0].parameters = (0.31,)
    assert fp(a) != fp(b)
    c = [gates.RY(0, 0.3), gates.CZ(0, 1), gates.Unitary(np.array([[0, 1], [1, 0]]), 2)]
    assert fp(a) != fp(c)
    assert fp(a) != fp([gates.RY(0, 0.3), gates.CZ(1, 0), gates.Unitary(np.eye(2), 2)])
    fused = circuits.variational(6).fuse(2)
    assert fp(fused.queue) == fp(circuits.variational(6).fuse(2).queue)
    assert fp(fused.queue) != fp(circuits.variational(6, seed=5).fuse(2).queue)
