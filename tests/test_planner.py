"""CPU tests of the pass planner (host logic only): partition + diagonal merging must be a
re-ordering of memory traffic, not of the circuit's meaning."""

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


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("n,tile_bits,run_bits", [(6, 4, 2), (8, 5, 3), (9, 6, 2), (10, 10, 5), (7, 12, 5)])
def test_plan_matches_gate_by_gate(n, tile_bits, run_bits, seed):
    glist = random_circuit_gates(n, 60, seed)
    st = R.random_state(n, "complex128", seed)
    plan = planner.plan_queue(glist, n, MATS, tile_bits, run_bits, max_diag_bits=4 + seed % 3)
    got = plan_interp.run_plan(st.copy(), plan, n, _raw(n))
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
        for op in seg[2]:
            assert len(set(op.targets) | set(op.controls)) == len(op.targets) + len(op.controls)
            if op.kind == "dense":
                assert set(op.targets) <= set(local) and len(op.targets) in (1, 2)
            else:
                assert len(op.targets) <= 10


def test_qft_needs_few_passes_and_merges_its_phase_ladders():
    n = 33
    plan = planner.plan_queue(circuits.qft(n).queue, n, MATS, 12, 5)
    assert all(seg[0] == "pass" for seg in plan)
    assert len(plan) <= 9                      # 577 gates, 577 passes gate by gate
    nops = sum(len(seg[2]) for seg in plan)
    assert nops < 577 // 3                     # 528 CU1 merged into far fewer tables


def test_qft_plan_small_matches_analytic():
    n = 10
    plan = planner.plan_queue(circuits.qft(n).queue, n, MATS, 6, 3)
    st = np.zeros(1 << n, dtype=np.complex128)
    st[0] = 1
    got = plan_interp.run_plan(st, plan, n, _raw(n))
    np.testing.assert_allclose(got, np.full(1 << n, 2.0 ** (-n / 2)), rtol=0, atol=1e-14)


def test_diagonal_controls_are_split_off():
    n = 8
    glist = [gates.CU1(1, 0, 0.3), gates.CU1(2, 0, 0.5), gates.CU1(3, 0, 0.7)]
    plan = planner.plan_queue(glist, n, MATS, 5, 3)
    assert len(plan) == 1 and len(plan[0][2]) == 1
    op = plan[0][2][0]
    assert op.kind == "diag" and op.controls == (n - 1,) and len(op.targets) == 3
