"""CPU tests of the HOST side of the pass kernel: planner + C++ encoder (`qj_program_encode`, the
same code path as `qj_program_create` minus the upload) interpreted by tests/pass_emulator.py,
against gate-by-gate application.  No GPU needed: the image format, the thread -> amplitude maps,
payloads, predicates and phase factors are checked here; the kernel itself by the -m gpu tests."""

import numpy as np
import pytest

from qibojit_b200 import circuits, gates
from tests import pass_emulator as E
from tests import refdispatch as R
from tests.circuits_random import random_circuit_gates


def _run(glist, st, n, dtype, **kw):
    state = st.astype(np.complex128).copy()
    nimages = 0
    for kind, item in E.encode(glist, n, dtype, **kw):
        if kind == "image":
            E.run_image(item, state)
            nimages += 1
        else:
            state = R.reference_run(state, [item], n)
    return state, nimages


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("seed", range(3))
@pytest.mark.parametrize("n,tile_bits,run_bits", [(6, 6, 2), (8, 6, 3), (9, 7, 1), (10, 10, 5), (11, 9, 4), (12, 11, 4)])
def test_encoded_program_matches_gate_by_gate(n, tile_bits, run_bits, seed, dtype):
    glist = random_circuit_gates(n, 60, seed + 10 * n)
    st = R.random_state(n, "complex128", seed)
    got, nimages = _run(glist, st, n, dtype, tile_bits=tile_bits, run_bits=run_bits, max_diag_bits=4 + 2 * (seed % 4))
    ref = R.reference_run(st, glist, n)
    assert nimages >= 1
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12 if dtype == "complex128" else 3e-5)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("name", ["qft", "variational", "supremacy", "quantum_volume"])
def test_encoded_benchmark_circuits(name, dtype):
    n = 12
    kw = {"depth": 3} if name in ("supremacy", "quantum_volume") else {}
    c = getattr(circuits, name)(n, **kw)
    st = np.zeros(1 << n, dtype=np.complex128)
    st[0] = 1
    ref = R.reference_run(st, c.queue, n)
    for zero_state in (False, True):
        got, _ = _run(c.queue, st, n, dtype, zero_state=zero_state)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12 if dtype == "complex128" else 3e-5)


def test_encoder_rejects_malformed_programs():
    """Error behaviour of the entry point (QJ_ERR_INVALID -> ValueError), without a device."""
    import ctypes

    from qibojit_b200 import _capi, planner

    lib = _capi.load()
    passes = np.zeros(1, dtype=planner.PASS_DTYPE)
    passes[0]["nlocal"] = 6
    passes[0]["local_bits"][:6] = [1, 2, 3, 4, 5, 6]          # the tile must contain index bit 0
    img = ctypes.c_void_p()
    rc = lib.qj_program_encode(_capi.QJ_C128, 8, passes.ctypes.data, 1, None, 0, None, 0, None, 0, ctypes.byref(img))
    with pytest.raises(ValueError):
        _capi.check(rc)
    rc = lib.qj_program_encode(7, 8, passes.ctypes.data, 1, None, 0, None, 0, None, 0, ctypes.byref(img))
    with pytest.raises(ValueError):
        _capi.check(rc)


def test_controlled_and_diagonal_gates_in_every_placement():
    """Controls / phase-table bits inside the registers, among the thread bits and outside the tile."""
    n = 10
    rng = np.random.default_rng(2)
    u = np.linalg.qr(rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)))[0]
    glist = [gates.H(q) for q in range(n)]
    for c, t in [(0, 9), (9, 0), (4, 5), (1, 8), (7, 2)]:
        glist += [gates.Unitary(u, t).controlled_by(c), gates.CU1(c, t, 0.37), gates.CNOT(c, t),
                  gates.CRZ(c, t, 0.21), gates.RZZ(c, t, 0.4)]
    glist += [gates.TOFFOLI(0, 5, 9), gates.CCZ(1, 4, 8), gates.SWAP(2, 7).controlled_by(9), gates.fSim(3, 6, 0.2, 0.5)]
    st = R.random_state(n, "complex128", 4)
    ref = R.reference_run(st, glist, n)
    for dtype in ("complex128", "complex64"):
        for tile_bits, run_bits in [(6, 2), (7, 3), (9, 4)]:
            got, _ = _run(glist, st, n, dtype, tile_bits=tile_bits, run_bits=run_bits)
            np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12 if dtype == "complex128" else 3e-5)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n,tile_bits,run_bits,count", [(8, 8, 3, 4000), (10, 7, 3, 6000), (13, 11, 4, 400), (14, 13, 5, 400)])
def test_long_programs_split_into_several_launches(n, tile_bits, run_bits, count, dtype):
    """Passes whose image exceeds the shared-memory budget are split into several launches: the
    outer slots and per-tile factor indices of the moved round are renumbered."""
    glist = [g for g in random_circuit_gates(n, count, 3) if len(g.target_qubits) <= 2]   # no raw gates: long passes
    if n == 10:
        # dense gates on the qubits of the low 7 index bits only, diagonal gates anywhere (outer
        # bits and outer controls included): the whole list is ONE pass with thousands of ops
        rng = np.random.default_rng(7)
        low = list(range(n - 7, n))
        glist = []
        for _ in range(count // 2):
            a, b = (int(v) for v in rng.choice(low, size=2, replace=False))
            c, d = (int(v) for v in rng.choice(n, size=2, replace=False))
            kind = int(rng.integers(0, 6))
            glist.append([gates.RY(a, 0.3), gates.RX(b, 0.7), gates.H(a), gates.CNOT(c, a) if c != a else gates.H(a),
                          gates.Unitary(np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))[0], a, b),
                          gates.SWAP(a, b)][kind])
            glist.append([gates.CU1(c, d, 0.11), gates.CZ(c, d), gates.RZZ(c, d, 0.2), gates.RZ(c, 0.4),
                          gates.CRZ(c, d, 0.3), gates.Z(d)][int(rng.integers(0, 6))])
    st = R.random_state(n, "complex128", 1)
    segs = E.encode(glist, n, dtype, tile_bits=tile_bits, run_bits=run_bits)
    assert all(kind == "image" for kind, _ in segs)
    if n <= 10:   # more launches than passes: at least one pass did not fit one program image
        import ctypes

        b = E.EncoderBackend(dtype)
        from qibojit_b200 import planner

        prog = planner.Program(b, glist, n, dtype=dtype, tile_bits=tile_bits, run_bits=run_bits)
        npasses = len(prog.passes)
        prog.segments = []
        assert sum(len(item.launches) for _, item in segs) > npasses
    state = st.copy()
    for _, item in segs:
        E.run_image(item, state)
    ref = R.reference_run(st, glist, n)
    np.testing.assert_allclose(state, ref, rtol=0, atol=1e-11 if dtype == "complex128" else 5e-4)


def test_default_geometry_images_fit_the_kernel_limits():
    for dtype, n in (("complex128", 16), ("complex64", 17)):
        for kind, item in E.encode(circuits.qft(n).queue + circuits.supremacy(n, depth=3).queue, n, dtype):
            assert kind == "image"
            for L in item.launches:
                assert L["smem"] <= 200 << 10 and L["blob_units"] <= 2560 and L["nh"] <= 8


def test_random_geometries_fuzz():
    """150 random (n, dtype, tile, run, table size, circuit, zero-state) combinations through
    planner + encoder + emulator (a longer run of the same loop found the r = T planner hang)."""
    rng = np.random.default_rng(2024)
    for _ in range(150):
        n = int(rng.integers(6, 11))
        dtype = ["complex128", "complex64"][int(rng.integers(0, 2))]
        T = int(rng.integers(6, min(n, 12 if dtype == "complex128" else 13) + 1))
        r = int(rng.integers(1, T + 1))
        mdb = int(rng.integers(2, 13))
        seed = int(rng.integers(0, 1 << 30))
        zero = bool(rng.integers(0, 2))
        glist = random_circuit_gates(n, int(rng.integers(5, 50)), seed)
        if rng.random() < 0.5:
            glist = [g for g in glist if len(g.target_qubits) <= 2]
        if rng.random() < 0.5:
            a, b = (int(v) for v in rng.choice(n, size=2, replace=False))
            glist.insert(int(rng.integers(0, len(glist) + 1)), gates.SWAP(a, b))
        st = np.zeros(1 << n, dtype=np.complex128)
        st[0] = 1
        if not zero:
            st = R.random_state(n, "complex128", seed % 1000)
        got, _ = _run(glist, st, n, dtype, tile_bits=T, run_bits=r, max_diag_bits=mdb, zero_state=zero)
        np.testing.assert_allclose(got, R.reference_run(st, glist, n), rtol=0,
                                   atol=1e-11 if dtype == "complex128" else 2e-4,
                                   err_msg=str(dict(n=n, dtype=dtype, T=T, r=r, mdb=mdb, seed=seed, zero=zero)))


def test_tile_without_free_high_bits_does_not_hang():
    """run_bits == tile_bits < nqubits used to loop forever in planner.partition."""
    glist = random_circuit_gates(8, 14, 334108495)
    st = np.zeros(1 << 8, dtype=np.complex128)
    st[0] = 1
    got, _ = _run(glist, st, 8, "complex128", tile_bits=6, run_bits=6, max_diag_bits=3, zero_state=True)
    np.testing.assert_allclose(got, R.reference_run(st, glist, 8), rtol=0, atol=1e-12)


def test_whole_gate_table_fuzz():
    """Every gate class of qibojit_b200.gates (with extra controls) through planner + encoder +
    emulator, and through the distributed planner with all ranks in one process."""
    from tests.circuits_random import random_gate_any
    from tests.virtual_ranks import run_virtual

    rng = np.random.default_rng(77)
    for case in range(80):
        n = int(rng.integers(7, 10))
        dtype = ["complex128", "complex64"][case % 2]
        glist = [random_gate_any(n, rng) for _ in range(int(rng.integers(5, 50)))]
        st = np.zeros(1 << n, dtype=np.complex128)
        st[0] = 1
        ref = R.reference_run(st, glist, n)
        atol = 1e-11 if dtype == "complex128" else 2e-4
        T = int(rng.integers(6, n + 1))
        got, _ = _run(glist, st, n, dtype, tile_bits=T, run_bits=int(rng.integers(1, T + 1)),
                      max_diag_bits=int(rng.integers(2, 13)), zero_state=bool(case % 3 == 0))
        np.testing.assert_allclose(got, ref, rtol=0, atol=atol, err_msg=f"case {case}")
        got, _ = run_virtual(glist, n, [2, 4, 8][case % 3], dtype)
        np.testing.assert_allclose(got, ref, rtol=0, atol=atol, err_msg=f"case {case} (distributed)")


def test_compile_circuit_cache_is_replaced_when_parameters_change():
    """B200Backend.compile_circuit on a device-free stand-in: same parameters -> the cached program,
    new parameters -> a new program and the old one closed (no pile-up in a variational loop)."""
    from qibojit_b200.backends.b200 import B200Backend

    b = E.EncoderBackend("complex128")
    b._device_index = 0
    b.circuit_fingerprint = B200Backend.circuit_fingerprint
    c = circuits.variational(8)
    p1 = B200Backend.compile_circuit(b, c, zero_state=True)
    assert B200Backend.compile_circuit(b, c, zero_state=True) is p1
    p0 = B200Backend.compile_circuit(b, c)                       # other options: its own entry
    assert p0 is not p1 and len(c.__dict__["_qj_programs"]) == 2
    c.queue[0].parameters = (0.123,)
    p2 = B200Backend.compile_circuit(b, c, zero_state=True)
    assert p2 is not p1 and p1.segments == [] and len(c.__dict__["_qj_programs"]) == 2
    assert B200Backend.compile_circuit(b, c, zero_state=True) is p2
