"""-m gpu parity tests of the multi-gate pass kernel (`qj_program_*`, block_kernels.cu) against
gate-by-gate application: the planner may only change the order of memory traffic."""

import numpy as np
import pytest

from qibojit_b200 import circuits, gates, planner
from tests import refdispatch as R
from tests.circuits_random import random_circuit_gates
from tests.gpu_utils import ATOL, backend

pytestmark = pytest.mark.gpu


def _run(b, glist, st, n, dtype, **kw):
    b.set_dtype(dtype)
    try:
        prog = planner.Program(b, glist, n, dtype=dtype, **kw)
        d = b.cast(st, dtype=dtype, copy=True)
        d = prog.run(d)
        out = b.to_numpy(d)
        stats = prog.stats()
        prog.close()
    finally:
        b.set_dtype("complex128")
    return out, stats


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("n,tile_bits,run_bits", [(4, 4, 2), (6, 4, 2), (8, 5, 3), (9, 6, 1), (10, 10, 5),
                                                  (7, 12, 5), (14, 12, 5), (16, 13, 6), (13, 11, 3)])
def test_program_matches_gate_by_gate(n, tile_bits, run_bits, seed, dtype):
    b = backend()
    glist = random_circuit_gates(n, 80, seed + 10 * n)
    st = R.random_state(n, dtype, seed)
    got, stats = _run(b, glist, st, n, dtype, tile_bits=tile_bits, run_bits=run_bits,
                      max_diag_bits=4 + 2 * (seed % 4))
    ref = R.reference_run(st, glist, n)
    np.testing.assert_allclose(got, ref, rtol=0, atol=ATOL[dtype] * 20 if dtype == "complex64" else 1e-12)
    assert stats["launches"] >= 1


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_program_matches_per_gate_kernels_n20(dtype):
    """Same circuit through the per-gate kernels (already pinned to the oracle) and through
    the pass kernel at the default tile geometry."""
    b = backend()
    n = 20
    glist = random_circuit_gates(n, 120, 5)
    st = R.random_state(n, dtype, 3)
    got, _ = _run(b, glist, st, n, dtype)
    b.set_dtype(dtype)
    try:
        d = b.cast(st, dtype=dtype, copy=True)
        for g in glist:
            d = b.apply_gate(g, d, n)
        ref = b.to_numpy(d)
    finally:
        b.set_dtype("complex128")
    np.testing.assert_allclose(got, ref, rtol=0, atol=ATOL[dtype])


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [12, 20, 26])
def test_qft_program_analytic(n, dtype):
    """QFT|0..0> = 2^(-n/2) everywhere."""
    b = backend()
    st = np.zeros(1 << n, dtype=dtype)
    st[0] = 1
    got, stats = _run(b, circuits.qft(n).queue, st, n, dtype)
    np.testing.assert_allclose(got, np.full(1 << n, 2.0 ** (-n / 2)), rtol=0,
                               atol=1e-13 if dtype == "complex128" else 1e-6)
    assert stats["launches"] < n


@pytest.mark.parametrize("workload", ["variational", "supremacy", "quantum_volume"])
def test_benchmark_circuits_program_vs_per_gate(workload):
    b = backend()
    n = 18
    c = getattr(circuits, workload)(n)
    st = np.zeros(1 << n, dtype=np.complex128)
    st[0] = 1
    got, _ = _run(b, c.queue, st, n, "complex128")
    d = b.cast(st, dtype="complex128", copy=True)
    for g in c.queue:
        d = b.apply_gate(g, d, n)
    np.testing.assert_allclose(got, b.to_numpy(d), rtol=0, atol=1e-12)


def test_program_rejects_nonlocal_dense_target():
    import ctypes

    from qibojit_b200 import _capi

    b = backend()
    ops = np.zeros(1, dtype=planner.OP_DTYPE)
    ops[0]["kind"] = _capi.QJ_OPK_DENSE1
    ops[0]["ntargets"] = 1
    ops[0]["targets"][0] = 9
    passes = np.zeros(1, dtype=planner.PASS_DTYPE)
    passes[0]["nlocal"] = 4
    passes[0]["nops"] = 1
    passes[0]["local_bits"][:4] = [0, 1, 2, 3]
    data = np.eye(2, dtype=np.complex128).reshape(-1)
    out = ctypes.c_void_p()
    rc = b._lib.qj_program_create(b._handle(), _capi.QJ_C128, 10, passes.ctypes.data, 1,
                                  ops.ctypes.data, 1, data.ctypes.data, 4, ctypes.byref(out))
    assert rc == _capi.QJ_ERR_INVALID
    assert b"not a local bit" in b._lib.qj_last_error()
