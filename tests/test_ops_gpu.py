"""-m gpu parity tests of state preparation, collapse, probabilities and sampling
(CUDA path through the C ABI vs golden vectors of the reference's ops.py and the oracle)."""

import numpy as np
import pytest

from oracle import oracle as O
from tests import cases, goldenio
from tests import refdispatch as R
from tests.gpu_utils import backend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", cases.DTYPES)
@pytest.mark.parametrize("nqubits", [1, 2, 7, 20])
def test_zero_state(nqubits, dtype, golden_ops):
    # /root/reference/tests/test_ops.py:12-22
    b = backend()
    st = b.to_numpy(b.zero_state(nqubits, dtype=dtype))
    assert st.dtype == np.dtype(dtype)
    expect = np.zeros(1 << nqubits, dtype=dtype)
    expect[0] = 1
    np.testing.assert_array_equal(st, expect)
    if nqubits == 7:
        np.testing.assert_array_equal(st, golden_ops[f"init|{dtype}"])


def test_zero_density_matrix():
    b = backend()
    rho = b.to_numpy(b.zero_state(3, density_matrix=True))
    expect = np.zeros((8, 8), dtype=np.complex128)
    expect[0, 0] = 1
    np.testing.assert_array_equal(rho, expect)


@pytest.mark.parametrize("dtype", cases.DTYPES)
def test_collapse_golden(dtype, golden_ops):
    # /root/reference/tests/test_ops.py:52-83 + golden vectors of ops.py:47-79
    b = backend()
    seed = 1000
    for nq, meas, res in cases.COLLAPSE:
        for normalize in (True, False):
            seed += 1
            st = R.random_state(nq, dtype, seed)
            shot = int("".join(map(str, res)), 2)
            d = b.cast(st, dtype=dtype, copy=True)
            out = b.to_numpy(b.collapse_state(d, sorted(meas), np.array([shot]), nq, normalize))
            ref = R.collapse(O, st.copy(), meas, shot, nq, normalize)
            atol = 1e-6 if dtype == "complex64" else 1e-13
            np.testing.assert_allclose(out, ref, rtol=0, atol=atol)
            key = goldenio.case_key("collapse", dtype, nq, meas, res, f"norm{int(normalize)}")
            gold = golden_ops[key]
            np.testing.assert_allclose(goldenio.subsample(out, seed) if nq > 12 else out, gold,
                                       rtol=0, atol=atol * 10)


@pytest.mark.parametrize("dtype", cases.DTYPES)
@pytest.mark.parametrize("nq,qubits", [(3, [0]), (4, [1, 3]), (5, [4, 0, 2]), (6, [0, 1, 2, 3, 4, 5]),
                                       (7, [6, 5]), (12, [11]), (12, [0, 11, 5]), (14, list(range(13))),
                                       (14, list(range(14))), (14, [13, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]),
                                       (20, [3, 17]), (20, list(range(4, 20)))])
def test_probabilities(nq, qubits, dtype):
    b = backend()
    st = R.random_state(nq, dtype, 11)
    ref = O.calculate_probabilities(st, qubits, nq)
    out = b.to_numpy(b.calculate_probabilities(b.cast(st, dtype=dtype), qubits, nq))
    assert out.dtype == ref.dtype
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-6 if dtype == "complex64" else 1e-14)
    assert abs(out.sum() - 1) < (1e-4 if dtype == "complex64" else 1e-12)


def test_norm():
    b = backend()
    st = R.random_state(16, "complex128", 5) * 3.0
    assert abs(b.calculate_norm(b.cast(st)) - 3.0) < 1e-12


def test_sampler_reference_known_answer():
    # /root/reference/tests/test_ops.py:236-256: bit-exact RNG stream
    import torch

    b = backend()
    target = [72, 65, 63, 54, 57, 55, 67, 50, 53, 67, 69, 68, 64, 68, 66, 62]
    for realtype in (torch.float32, torch.float64):
        probs = torch.ones(16, dtype=realtype, device=b.torch_device) / 16
        freq = torch.zeros(16, dtype=torch.int64, device=b.torch_device)
        freq = b.measure_frequencies_op(freq, probs, nshots=1000, nqubits=4, seed=1234, nthreads=4)
        assert int(freq.sum()) == 1000
        np.testing.assert_array_equal(freq.cpu().numpy(), np.array(target))


def test_sampler_golden(golden_ops):
    """Metropolis sampler vs the reference's numba output, bit for bit (ops.py:86-108)."""
    import torch

    b = backend()
    for key in golden_ops.files:
        parts = key.split("|")
        if parts[0] != "freq":
            continue
        realtype, nq, nshots, seedv, nthreads = (parts[1], int(parts[2][1:]), int(parts[3][1:]),
                                                 int(parts[4][4:]), int(parts[5][1:]))
        rng = np.random.default_rng(nq * 131 + nshots)
        probs = rng.random(1 << nq)
        if nq == 4:
            probs = np.ones(16)
        probs = (probs / probs.sum()).astype(realtype)
        dprobs = torch.as_tensor(probs, device=b.torch_device)
        freq = torch.zeros(1 << nq, dtype=torch.int64, device=b.torch_device)
        b.measure_frequencies_op(freq, dprobs, nshots, nq, seedv, nthreads)
        np.testing.assert_array_equal(freq.cpu().numpy(), golden_ops[key])


def test_sampler_long_chain_matches_oracle():
    """Several MT19937 re-twists per chain (>> 624 draws) and a non power-of-two ratio test."""
    import torch

    b = backend()
    nq = 12
    rng = np.random.default_rng(5)
    probs = rng.random(1 << nq) ** 3
    probs /= probs.sum()
    ref = O.measure_frequencies(np.zeros(1 << nq, dtype=np.int64), probs, 200000, nq, 4242, 6)
    freq = torch.zeros(1 << nq, dtype=torch.int64, device=b.torch_device)
    b.measure_frequencies_op(freq, torch.as_tensor(probs, device=b.torch_device), 200000, nq, 4242, 6)
    np.testing.assert_array_equal(freq.cpu().numpy(), ref)


@pytest.mark.parametrize("nshots", [1000, 200000])
def test_sample_frequencies_sparse_support(nshots):
    # /root/reference/tests/test_ops.py:259-275
    import itertools

    b = backend()
    for nonzero in list(itertools.combinations(range(8), r=2))[:6] + [(3,), (0, 1, 2, 7)]:
        probs = np.zeros(8)
        probs[list(nonzero)] = 1
        probs /= probs.sum()
        freqs = b.sample_frequencies(b.cast(probs, dtype="float64"), nshots)
        assert sum(freqs.values()) == nshots
        assert set(freqs) <= set(nonzero)


def test_sample_shots_matches_numpy_choice():
    """Low-shot path: numpy's legacy choice(p=...) = cumsum / searchsorted(side='right')."""
    b = backend()
    rng = np.random.default_rng(8)
    probs = rng.random(1 << 10)
    probs /= probs.sum()
    b.set_seed(77)
    shots = b.to_numpy(b.sample_shots(b.cast(probs, dtype="float64"), 5000))
    np.random.seed(77)
    expect = np.random.choice(len(probs), size=5000, p=probs)
    assert (shots == expect).mean() > 0.999  # cumsum association may flip a boundary ulp


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_swap_pack_bits_roundtrip(dtype):
    """qj_swap_pack_bits gathers the sub-block selected by several index bits (the piece that goes
    to one peer of a multi-qubit exchange), qj_swap_unpack_bits scatters into the same slots."""
    import torch

    from qibojit_b200 import _capi

    b = backend()
    nlocal = 12
    rng = np.random.default_rng(5)
    host = (rng.standard_normal(1 << nlocal) + 1j * rng.standard_normal(1 << nlocal)).astype(dtype)
    shard = b.cast(host, dtype=dtype, copy=True)
    tag = b._tag(shard)
    for bits, value in [([3, 7], 2), ([1, 5, 11], 5), ([2], 1), ([4, 6, 8, 10], 9)]:
        k = len(bits)
        idx = np.arange(1 << nlocal)
        field = np.zeros_like(idx)
        for i, l in enumerate(bits):
            field |= ((idx >> l) & 1) << i
        sel = idx[field == value]
        sub = 1 << (nlocal - k)
        barr = np.asarray(bits, dtype=np.int32)
        buf = torch.empty(sub, dtype=shard.dtype, device=shard.device)
        half = sub // 2
        for c0, n in ((0, half), (half, sub - half)):       # two chunks
            _capi.check(b._lib.qj_swap_pack_bits(b._handle(), shard.data_ptr(), buf[c0:].data_ptr(), tag, nlocal,
                                                 barr.ctypes.data, k, value, c0, n))
        np.testing.assert_array_equal(b.to_numpy(buf), host[sel])
        repl = (rng.standard_normal(sub) + 1j * rng.standard_normal(sub)).astype(dtype)
        src = b.cast(repl, dtype=dtype, copy=True)
        _capi.check(b._lib.qj_swap_unpack_bits(b._handle(), shard.data_ptr(), src.data_ptr(), tag, nlocal,
                                               barr.ctypes.data, k, value, 0, sub))
        host[sel] = repl
        np.testing.assert_array_equal(b.to_numpy(shard), host)
    with pytest.raises(ValueError):
        barr = np.asarray([5, 3], dtype=np.int32)           # not ascending
        _capi.check(b._lib.qj_swap_pack_bits(b._handle(), shard.data_ptr(), shard.data_ptr(), tag, nlocal,
                                             barr.ctypes.data, 2, 0, 0, 2))


SWAP_PIECES_CASES = [(3, 0), (3, 2), (5, 1), (9, 0), (9, 4), (9, 8)]     # (nlocal, new_global) of the golden generator


def _swap_pieces_inputs(dtype):
    """The golden generator's inputs (tests/golden/make_golden.py: seeds continue after the collapse cases)."""
    seed = 1000 + 2 * len(cases.COLLAPSE)
    out = []
    for nlocal, new_global in SWAP_PIECES_CASES:
        seed += 1
        full = R.random_state(nlocal + 1, dtype, seed)
        out.append((nlocal, new_global, full[: 1 << nlocal].copy(), full[1 << nlocal:].copy()))
    return out


@pytest.mark.parametrize("dtype", cases.DTYPES)
@pytest.mark.parametrize("transport", ["pieces_peer", "bits_peer", "pack_unpack"])
def test_swap_pieces_golden(transport, dtype, golden_ops):
    """ops.swap_pieces (ops.py:131-137; pinned by the reference at tests/test_ops.py:168-233) through the
    three device transports of a global<->local qubit swap: the in-place peer kernels
    (`qj_swap_pieces_peer`, `qj_swap_bits_peer`; both pieces on this device play the two ranks) and
    the pack -> transfer -> unpack path of the NCCL transport (`qj_swap_pack` / `qj_swap_unpack`).
    Bit-exact against the reference's numba output."""
    import torch

    from qibojit_b200 import _capi

    b = backend()
    lib, h = b._lib, b._handle()
    for nlocal, new_global, h0, h1 in _swap_pieces_inputs(dtype):
        m = nlocal - new_global - 1
        p0, p1 = b.cast(h0, dtype=dtype, copy=True), b.cast(h1, dtype=dtype, copy=True)
        tag = b._tag(p0)
        gold = golden_ops[f"swap_pieces|{dtype}|l{nlocal}|g{new_global}"]
        if dtype == "complex64" and m == 0:
            # 16-byte granularity: the planner never picks index bit 0 of a complex64 shard
            with pytest.raises(NotImplementedError):
                _capi.check(lib.qj_swap_pieces_peer(h, p0.data_ptr(), p1.data_ptr(), tag, nlocal, m, 0))
            continue
        if transport == "pieces_peer":
            _capi.check(lib.qj_swap_pieces_peer(h, p0.data_ptr(), p1.data_ptr(), tag, nlocal, m, 0))
            _capi.check(lib.qj_swap_pieces_peer(h, p1.data_ptr(), p0.data_ptr(), tag, nlocal, m, 1))
        elif transport == "bits_peer":
            bits = np.asarray([m], dtype=np.int32)
            _capi.check(lib.qj_swap_bits_peer(h, p0.data_ptr(), p1.data_ptr(), tag, nlocal, bits.ctypes.data, 1, 1, 0, 0, 2))
            _capi.check(lib.qj_swap_bits_peer(h, p1.data_ptr(), p0.data_ptr(), tag, nlocal, bits.ctypes.data, 1, 0, 1, 1, 2))
        else:
            half = 1 << (nlocal - 1)
            s0 = torch.empty(half, dtype=p0.dtype, device=p0.device)
            s1 = torch.empty(half, dtype=p0.dtype, device=p0.device)
            cut = (half // 2) & ~1                                    # two chunks, vector aligned
            for c0, n in ((0, cut), (cut, half - cut)):
                if n == 0:
                    continue
                _capi.check(lib.qj_swap_pack(h, p0.data_ptr(), s0[c0:].data_ptr(), tag, nlocal, m, 0, c0, n))
                _capi.check(lib.qj_swap_pack(h, p1.data_ptr(), s1[c0:].data_ptr(), tag, nlocal, m, 1, c0, n))
            for c0, n in ((0, cut), (cut, half - cut)):
                if n == 0:
                    continue
                _capi.check(lib.qj_swap_unpack(h, p0.data_ptr(), s1[c0:].data_ptr(), tag, nlocal, m, 0, c0, n))
                _capi.check(lib.qj_swap_unpack(h, p1.data_ptr(), s0[c0:].data_ptr(), tag, nlocal, m, 1, c0, n))
        got = np.concatenate([b.to_numpy(p0), b.to_numpy(p1)])
        np.testing.assert_array_equal(got, gold, err_msg=f"{transport} nlocal={nlocal} new_global={new_global}")


@pytest.mark.parametrize("dtype", cases.DTYPES)
def test_swap_bits_peer_all_to_all(dtype):
    """The multi-qubit exchange over peer memory: 2^k 'ranks' (all on this device) swap sub-blocks
    pairwise; afterwards the k exchanged local bits and the k rank bits have traded places."""
    from qibojit_b200 import _capi

    b = backend()
    lib, h = b._lib, b._handle()
    nlocal, bits = 10, [3, 7, 9]
    k = len(bits)
    rng = np.random.default_rng(11)
    host = [(rng.standard_normal(1 << nlocal) + 1j * rng.standard_normal(1 << nlocal)).astype(dtype) for _ in range(1 << k)]
    dev = [b.cast(x, dtype=dtype, copy=True) for x in host]
    tag = b._tag(dev[0])
    barr = np.asarray(bits, dtype=np.int32)
    for r in range(1 << k):
        for d in range(1, 1 << k):
            a = r ^ d
            _capi.check(lib.qj_swap_bits_peer(h, dev[r].data_ptr(), dev[a].data_ptr(), tag, nlocal, barr.ctypes.data, k,
                                              a, r, 0 if r < a else 1, 2))
    idx = np.arange(1 << nlocal)
    field = np.zeros_like(idx)
    for i, l in enumerate(bits):
        field |= ((idx >> l) & 1) << i
    for r in range(1 << k):
        want = host[r].copy()
        for a in range(1 << k):
            if a != r:
                want[field == a] = host[a][field == r]
        np.testing.assert_array_equal(b.to_numpy(dev[r]), want)
