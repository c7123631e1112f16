"""CPU suite: pins the oracle (oracle/qj_oracle.c) against
 (1) golden vectors produced by the reference's own numba kernels (tests/golden/),
 (2) the reference's sampler known-answer vector (tests/test_ops.py:251-256),
 (3) an independent einsum oracle (the role qibo's NumpyBackend plays upstream)."""

import functools

import numpy as np
import pytest

from oracle import oracle as O
from tests import cases, goldenio
from tests import refdispatch as R

ATOL = {"complex64": 1e-6, "complex128": 1e-14}

ORACLE_DISPATCH = (
    functools.partial(R.one_qubit_base, O),
    functools.partial(R.two_qubit_base, O),
    functools.partial(R.multi_qubit_base, O),
)

GATE_CASES = list(goldenio.iter_gate_cases(cases, R))


@pytest.mark.parametrize("case", GATE_CASES, ids=[c[0] for c in GATE_CASES])
def test_oracle_matches_reference_golden(case, golden_gates):
    key, kind, dtype, nq, t, c, seed = case
    out = goldenio.run_gate_case(ORACLE_DISPATCH, R, kind, dtype, nq, t, c, seed)
    if kind == "multilarge":
        out = goldenio.subsample(out, seed)
        np.testing.assert_allclose(out, golden_gates[key], atol=1e-4 if dtype == "complex64" else 1e-11)
    else:
        np.testing.assert_allclose(out, golden_gates[key], rtol=0, atol=ATOL[dtype])


@pytest.mark.parametrize("case", [c for c in GATE_CASES if c[1] in ("gate", "two", "multi")],
                         ids=[c[0] for c in GATE_CASES if c[1] in ("gate", "two", "multi")])
def test_oracle_matches_einsum(case):
    key, kind, dtype, nq, t, c, seed = case
    tl = [t] if kind == "gate" else list(t)
    dim = 1 << len(tl)
    st = R.random_state(nq, dtype, seed)
    expect = R.einsum_apply(st, R.random_matrix(dim, dtype, seed), tl, c, nq)
    out = goldenio.run_gate_case(ORACLE_DISPATCH, R, kind, dtype, nq, t, c, seed)
    np.testing.assert_allclose(out, expect, atol=1e-4 if dtype == "complex64" else 1e-10)


def test_sampler_reference_known_answer():
    # /root/reference/tests/test_ops.py:236-256
    target = [72, 65, 63, 54, 57, 55, 67, 50, 53, 67, 69, 68, 64, 68, 66, 62]
    for realtype in ("float32", "float64"):
        for inttype in ("int32", "int64"):
            probs = np.ones(16, dtype=realtype) / 16
            freq = np.zeros(16, dtype=inttype)
            freq = O.measure_frequencies(freq, probs, nshots=1000, nqubits=4, seed=1234, nthreads=4)
            assert freq.sum() == 1000
            np.testing.assert_array_equal(freq, np.array(target, dtype=inttype))


def test_mt19937_matches_numpy_legacy_stream():
    rs = np.random.RandomState(1234)
    np.testing.assert_array_equal(O.mt_doubles(1234, 2000), rs.random_sample(2000))


def test_ops_golden(golden_ops):
    for key in golden_ops.files:
        parts = key.split("|")
        kind = parts[0]
        if kind == "init":
            st = np.empty(1 << 7, dtype=parts[1])
            np.testing.assert_array_equal(O.initial_state_vector(st), golden_ops[key])
        elif kind == "freq":
            realtype, nq, nshots, seedv, nthreads = (parts[1], int(parts[2][1:]), int(parts[3][1:]),
                                                     int(parts[4][4:]), int(parts[5][1:]))
            rng = np.random.default_rng(nq * 131 + nshots)
            probs = rng.random(1 << nq)
            if nq == 4:
                probs = np.ones(16)
            probs = (probs / probs.sum()).astype(realtype)
            freq = O.measure_frequencies(np.zeros(1 << nq, dtype=np.int64), probs, nshots, nq,
                                         seedv, nthreads)
            np.testing.assert_array_equal(freq, golden_ops[key])


def test_collapse_golden(golden_ops):
    for dtype in cases.DTYPES:
        seed = 1000
        for nq, meas, res in cases.COLLAPSE:
            for normalize in (True, False):
                seed += 1
                st = R.random_state(nq, dtype, seed)
                shot = int("".join(map(str, res)), 2)
                out = R.collapse(O, st, meas, shot, nq, normalize)
                if nq > 12:
                    out = goldenio.subsample(out, seed)
                key = goldenio.case_key("collapse", dtype, nq, meas, res, f"norm{int(normalize)}")
                np.testing.assert_allclose(out, golden_ops[key], rtol=0,
                                           atol=1e-6 if dtype == "complex64" else 1e-14)


def test_collapse_vs_slicing():
    # known-answer construction of /root/reference/tests/test_ops.py:64-83
    for dtype in cases.DTYPES:
        for nq, meas, res in cases.COLLAPSE:
            st = R.random_state(nq, dtype, 5)
            slicer = nq * [slice(None)]
            for t, r in zip(meas, res):
                slicer[t] = r
            init = np.reshape(np.copy(st), nq * (2,))
            target = np.zeros_like(init)
            target[tuple(slicer)] = init[tuple(slicer)]
            target = target.flatten()
            target = target / np.sqrt((np.abs(target) ** 2).sum())
            out = R.collapse(O, st, meas, int("".join(map(str, res)), 2), nq, True)
            np.testing.assert_allclose(out, target, atol=1e-6 if dtype == "complex64" else 1e-14)


def test_swap_pieces_and_transpose_golden(golden_ops):
    for key in golden_ops.files:
        parts = key.split("|")
        if parts[0] not in ("swap_pieces", "transpose"):
            continue
    for dtype in cases.DTYPES:
        seed = 1000 + 2 * len(cases.COLLAPSE)
        for nlocal, new_global in [(3, 0), (3, 2), (5, 1), (9, 0), (9, 4), (9, 8)]:
            seed += 1
            full = R.random_state(nlocal + 1, dtype, seed)
            p0, p1 = full[: 1 << nlocal].copy(), full[1 << nlocal:].copy()
            O.swap_pieces(p0, p1, new_global, nlocal)
            np.testing.assert_array_equal(np.concatenate([p0, p1]),
                                          golden_ops[f"swap_pieces|{dtype}|l{nlocal}|g{new_global}"])
        rng = np.random.default_rng(77)
        for nq, ndev in [(3, 2), (5, 4), (8, 8), (10, 2)]:
            seed += 1
            order = [int(v) for v in rng.permutation(nq)]
            full = R.random_state(nq, dtype, seed)
            pieces = [p.copy() for p in full.reshape(ndev, -1)]
            out = O.transpose_state(pieces, np.zeros_like(full), nq, order)
            key = f"transpose|{dtype}|n{nq}|d{ndev}|o{'_'.join(map(str, order))}"
            np.testing.assert_array_equal(out, golden_ops[key])
            # and the reference test's own known answer (tests/test_ops.py:150-165)
            np.testing.assert_array_equal(
                out, np.transpose(full.reshape(nq * (2,)), order).flatten())


def test_probabilities_definition():
    # qibo Backend.calculate_probabilities semantics (third-party; parity unpinned):
    # checked against the plain numpy definition.
    for dtype in cases.DTYPES:
        for nq, qubits in [(3, [0]), (4, [1, 3]), (5, [4, 0, 2]), (6, list(range(6))), (7, [6, 5])]:
            st = R.random_state(nq, dtype, 11)
            p = (np.abs(st.astype(np.complex128)) ** 2).reshape(nq * (2,))
            unmeasured = tuple(q for q in range(nq) if q not in qubits)
            p = p.sum(axis=unmeasured)
            order = np.argsort(np.argsort(qubits))  # axes currently sorted ascending
            p = np.transpose(p, [sorted(qubits).index(q) for q in qubits]).ravel()
            out = O.calculate_probabilities(st, qubits, nq)
            np.testing.assert_allclose(out, p, atol=1e-6 if dtype == "complex64" else 1e-14)
