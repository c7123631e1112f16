"""-m gpu parity tests of the multi-gate pass kernel (`qj_program_*`, pass_kernels.cu) against
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
@pytest.mark.parametrize("n,tile_bits,run_bits", [(4, 4, 2), (6, 6, 2), (8, 6, 3), (9, 7, 1), (10, 10, 5),
                                                  (7, 12, 5), (14, 12, 5), (16, 13, 6), (13, 11, 3),
                                                  (15, 12, 4), (12, 9, 6)])
def test_program_matches_gate_by_gate(n, tile_bits, run_bits, seed, dtype):
    b = backend()
    glist = random_circuit_gates(n, 80, seed + 10 * n)
    st = R.random_state(n, dtype, seed)
    got, stats = _run(b, glist, st, n, dtype, tile_bits=tile_bits, run_bits=run_bits,
                      max_diag_bits=4 + 2 * (seed % 4))
    ref = R.reference_run(st, glist, n)
    np.testing.assert_allclose(got, ref, rtol=0, atol=ATOL[dtype])
    assert stats["launches"] >= (1 if n >= planner.MIN_QUBITS else 0)


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


def _create(b, n, passes, rounds, ops, data):
    import ctypes

    from qibojit_b200 import _capi

    out = ctypes.c_void_p()
    rc = b._lib.qj_program_create(b._handle(), _capi.QJ_C128, n, passes.ctypes.data, len(passes),
                                  rounds.ctypes.data, len(rounds), ops.ctypes.data, len(ops),
                                  data.ctypes.data, data.size, ctypes.byref(out))
    return rc, out


def test_program_rejects_dense_target_outside_the_registers():
    from qibojit_b200 import _capi

    b = backend()
    ops = np.zeros(1, dtype=planner.OP_DTYPE)
    ops[0]["kind"] = _capi.QJ_OPK_DENSE1
    ops[0]["ntargets"] = 1
    rounds = np.zeros(1, dtype=planner.ROUND_DTYPE)
    rounds[0]["nreg"] = 4
    rounds[0]["nops"] = 1
    rounds[0]["reg_bits"][:4] = [0, 1, 2, 3]
    passes = np.zeros(1, dtype=planner.PASS_DTYPE)
    passes[0]["nlocal"] = 6
    passes[0]["nrounds"] = 1
    passes[0]["local_bits"][:6] = [0, 1, 2, 3, 4, 5]
    data = np.eye(2, dtype=np.complex128).reshape(-1)
    for target, msg in [(9, b"not a local bit"), (5, b"not a register bit")]:
        ops[0]["targets"][0] = target
        rc, _ = _create(b, 10, passes, rounds, ops, data)
        assert rc == _capi.QJ_ERR_INVALID
        assert msg in b._lib.qj_last_error()
    rounds[0]["reg_bits"][:4] = [0, 1, 2, 8]
    rc, _ = _create(b, 10, passes, rounds, ops, data)
    assert rc == _capi.QJ_ERR_INVALID and b"register bit is not a local bit" in b._lib.qj_last_error()


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", range(3))
def test_zero_state_program_relabels_swaps(seed, dtype):
    """Programs compiled for the |0...0> input drop the SWAP gates (planner.relabel_swaps_away):
    same final state as gate by gate, through execute_circuit and through Program directly."""
    from qibojit_b200.circuit import Circuit

    b = backend()
    n = 14
    rng = np.random.default_rng(seed)
    glist = []
    for g in random_circuit_gates(n, 60, seed + 7):
        if len(g.target_qubits) > 2:
            continue
        glist.append(g)
        if rng.random() < 0.2:
            a, c = rng.choice(n, size=2, replace=False)
            glist.append(gates.SWAP(int(a), int(c)))
    glist += [gates.SWAP(i, n - 1 - i) for i in range(n // 2)]
    st = np.zeros(1 << n, dtype=dtype)
    st[0] = 1
    ref = R.reference_run(st.astype(np.complex128), glist, n)
    got, _ = _run(b, glist, st, n, dtype, zero_state=True)
    atol = ATOL[dtype]
    np.testing.assert_allclose(got, ref, rtol=0, atol=atol)
    c = Circuit(n)
    c.add(glist)
    b.set_dtype(dtype)
    try:
        out = b.to_numpy(b.execute_circuit(c))
        np.testing.assert_allclose(out, ref, rtol=0, atol=atol)
        # a given initial state keeps the swaps
        rs = R.random_state(n, dtype, seed)
        out = b.to_numpy(b.execute_circuit(c, initial_state=rs))
        np.testing.assert_allclose(out, R.reference_run(rs, glist, n), rtol=0, atol=atol)
    finally:
        b.set_dtype("complex128")


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [8, 13, 17])
def test_program_from_zero_ignores_the_buffer(n, dtype):
    """`run(state, from_zero=True)`: |0...0> preparation (ops.py:14-18) fused into the first pass -- the
    first launch writes every amplitude without reading any, so garbage in the buffer must not matter;
    a program whose first segment is a raw gate falls back to the initial-state kernel."""
    import torch

    b = backend()
    b.set_dtype(dtype)
    try:
        for seed, head in ((1, []), (2, [gates.Unitary(R.random_unitary(8, 3), 0, 1, 2)])):     # raw 3-qubit gate first
            glist = head + random_circuit_gates(n, 50, seed + n)
            st = np.zeros(1 << n, dtype=dtype)
            st[0] = 1
            ref = R.reference_run(st, glist, n)
            prog = planner.Program(b, glist, n, dtype=dtype)
            garbage = torch.full((1 << n,), float("nan"), dtype=getattr(torch, dtype), device=b.torch_device)
            out = b.to_numpy(prog.run(garbage, from_zero=True))
            np.testing.assert_allclose(out, ref, rtol=0, atol=ATOL[dtype])
            want = b.to_numpy(prog.run(b.cast(st, dtype=dtype, copy=True)))
            np.testing.assert_array_equal(out, want)           # bit-identical to the explicit preparation
            prog.close()
        # an empty program still prepares the state
        prog = planner.Program(b, [], n, dtype=dtype)
        garbage = torch.full((1 << n,), float("nan"), dtype=getattr(torch, dtype), device=b.torch_device)
        out = b.to_numpy(prog.run(garbage, from_zero=True))
        assert out[0] == 1 and not np.any(out[1:])
    finally:
        b.set_dtype("complex128")


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("pieces", [1, 2, 8])
def test_launches_by_tile_range_and_out_of_place(dtype, pieces):
    """`qj_program_run_tiles` over a partition of the tiles equals the whole launch, and
    `qj_program_run_tiles_to` stores the same amplitudes into another buffer (the distributed layer
    pipelines the last pass before an exchange this way)."""
    import ctypes

    from qibojit_b200 import _capi

    b = backend()
    n = 18
    glist = random_circuit_gates(n, 60, 11, with_raw=False)
    st = R.random_state(n, dtype, 4)
    b.set_dtype(dtype)
    try:
        prog = planner.Program(b, glist, n, dtype=dtype)
        whole = b.to_numpy(prog.run(b.cast(st, dtype=dtype, copy=True)))
        assert [s[0] for s in prog.segments] == ["program"]
        handle, h = prog.segments[0][1], b._handle()
        nl = ctypes.c_int64()
        _capi.check(b._lib.qj_program_stats(handle, ctypes.byref(nl), None, None))
        last = int(nl.value) - 1
        geom = (ctypes.c_int64 * 12)()
        _capi.check(b._lib.qj_program_launch_geometry(handle, last, geom))
        ntiles = int(geom[3])
        assert ntiles % pieces == 0
        d = b.cast(st, dtype=dtype, copy=True)
        other = b.cast(np.zeros_like(st), dtype=dtype, copy=True)
        if last > 0:
            _capi.check(b._lib.qj_program_run_ex(h, handle, d.data_ptr(), 0, last, 0))
        step = ntiles // pieces
        # even pieces in place, odd pieces into the other buffer
        for i in range(pieces):
            if i % 2 == 0:
                _capi.check(b._lib.qj_program_run_tiles(h, handle, d.data_ptr(), last, i * step, step))
            else:
                _capi.check(b._lib.qj_program_run_tiles_to(h, handle, d.data_ptr(), ctypes.c_void_p(other.data_ptr()),
                                                           last, i * step, step))
        got_d, got_o = b.to_numpy(d), b.to_numpy(other)
        # tiles are numbered by the index bits outside the tile, least significant first: piece i
        # holds the amplitudes whose outside bits spell a tile number in its range
        T_r, hib = int(geom[1]), [int(v) for v in geom[4:12] if v >= 0]
        outside = [q for q in range(T_r, n) if q not in hib]
        idx = np.arange(1 << n)
        tile = np.zeros(1 << n, dtype=np.int64)
        for j, q in enumerate(outside):
            tile |= ((idx >> q) & 1) << j
        piece_of = tile // step
        in_place = piece_of % 2 == 0
        np.testing.assert_array_equal(got_d[in_place], whole[in_place])
        np.testing.assert_array_equal(got_o[~in_place], whole[~in_place])
        assert not got_o[in_place].any()
        # out of range
        assert b._lib.qj_program_run_tiles(h, handle, d.data_ptr(), last, ntiles, 1) != 0
        prog.close()
    finally:
        b.set_dtype("complex128")


def test_peer_handshake_and_async_copy_on_one_device():
    """`qj_peer_handshake` against this device's own flag words (the peer is this rank: the epoch it
    waits for is the one it publishes) and `qj_copy_async` between two buffers."""
    import ctypes

    import torch

    from qibojit_b200 import _capi

    b = backend()
    h = b._handle()
    flags = torch.zeros(64, dtype=torch.int32, device=b.torch_device)
    for epoch in (1, 2, 7):
        slots = (ctypes.c_void_p * 2)(flags.data_ptr() + 4 * 3, flags.data_ptr() + 4 * 5)
        src = (ctypes.c_int32 * 2)(3, 5)
        eps = (ctypes.c_uint32 * 2)(epoch, epoch + 100)
        _capi.check(b._lib.qj_peer_handshake(h, ctypes.c_void_p(flags.data_ptr()), slots, src, eps, 2,
                                             ctypes.c_double(5.0)))
        b.synchronize()
        got = flags.cpu().numpy()
        assert got[3] == epoch and got[5] == epoch + 100 and got[[0, 1, 2, 4]].sum() == 0
    assert b._lib.qj_peer_handshake(h, ctypes.c_void_p(flags.data_ptr()), slots, src, eps, 0, ctypes.c_double(1.0)) != 0
    a = torch.arange(1 << 20, dtype=torch.float64, device=b.torch_device)
    c = torch.zeros_like(a)
    _capi.check(b._lib.qj_copy_async(h, ctypes.c_void_p(c.data_ptr()), ctypes.c_void_p(a.data_ptr()), a.numel() * 8))
    b.synchronize()
    assert torch.equal(a, c)
