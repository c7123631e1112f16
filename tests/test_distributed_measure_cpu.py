"""world_size-2/4 gloo tests (CPU ranks, oracle kernels) of the distributed state's measurement and
layout functions: `normalize_layout` / `to_tensor` (ops.transpose_state semantics, ops.py:112-124),
`collapse` (ops.py:47-79 on shards), `sample_frequencies` (bit-exact with ops.py:86-108 run on the
whole vector) and `execute_distributed_circuit(initial_state=...)` (gpu.py:646-739)."""

import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _circuit(n):
    from qibojit_b200 import circuits, gates
    from qibojit_b200.circuit import Circuit

    c = Circuit(n)
    c.add(circuits.qft(n).queue)                     # SWAPs -> relabellings: a permuted final map
    c.add([gates.RY(0, 0.3), gates.CNOT(0, n - 1), gates.SWAP(1, n - 2), gates.H(n - 1), gates.CU1(0, 2, 0.4),
           gates.RX(1, 0.7), gates.SWAP(0, 2)])
    return c


def _worker(rank, world, port, n, dtype, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch

        from oracle import oracle as O
        from qibojit_b200 import distributed as D
        from tests.oracle_backend import OracleBackend
        from tests.test_distributed_cpu import _reference_state

        O.set_threads(1)
        torch.set_num_threads(1)
        b = OracleBackend(dtype)
        circuit = _circuit(n)
        ref = _reference_state(circuit, dtype)
        out = {}

        # layout normalisation: the shard becomes the reference's piece, to_tensor the logical vector
        ds = D.DistributedState(b, n, comm=D.Comm(), dtype=dtype)
        ds.execute(circuit.queue)
        permuted = ds.bit_of != [n - 1 - k for k in range(n)]
        full = ds.to_tensor().numpy().copy()
        assert ds.bit_of == [n - 1 - k for k in range(n)]
        size = 1 << ds.nlocal
        piece = ds.to_pieces().numpy()
        out["to_tensor"] = (full, permuted, np.array_equal(piece, full[rank * size:(rank + 1) * size]))

        # sampling: Metropolis sampler on the gathered probabilities, same seed on every rank
        np.random.seed(7)
        freqs = ds.sample_frequencies(200000)
        out["freqs"] = dict(freqs)

        # collapse on (local, global, relabelled) qubits, normalised and not
        for tag, qubits, shot, normalize in [("c1", [0, n - 1], 2, True), ("c2", [1], 1, True),
                                             ("c3", [0, 2, 3], 5, False), ("c4", list(range(n)), 6, True)]:
            d2 = D.DistributedState(b, n, comm=D.Comm(), dtype=dtype)
            d2.execute(circuit.queue)
            d2.collapse(qubits, shot, normalize=normalize)
            out[tag] = d2.to_numpy_full()

        # initial state + execute_distributed_circuit
        rng = np.random.default_rng(5)
        init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        init = (init / np.linalg.norm(init)).astype(dtype)
        d3 = D.execute_distributed_circuit(b, circuit, initial_state=init, comm=D.Comm())
        out["init"] = (init, d3.to_tensor().numpy().copy())
        d4 = D.execute_distributed_circuit(b, circuit, initial_state=torch.from_numpy(init), comm=D.Comm())
        assert np.array_equal(d4.to_tensor().numpy(), out["init"][1])
        try:
            D.execute_distributed_circuit(b, circuit, initial_state="zeros", comm=D.Comm())
            raise AssertionError("a string is not an initial state")
        except TypeError:
            pass
        np.random.seed(11)
        d5, f5 = D.execute_distributed_circuit(b, circuit, nshots=150000, comm=D.Comm())
        out["nshots"] = dict(f5)
        if rank == 0:
            q.put((ref, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,dtype", [(2, 7, "complex128"), (4, 8, "complex64"), (8, 9, "complex128")])
def test_distributed_measurement_and_layout(world, n, dtype):
    from oracle import oracle as O
    from tests import refdispatch as R
    from tests.test_distributed_cpu import _reference_state

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, dtype, q)) for r in range(world)]
    for p in procs:
        p.start()
    ref, out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    atol = 1e-5 if dtype == "complex64" else 1e-12
    full, permuted, piece_ok = out["to_tensor"]
    assert permuted, "the circuit must leave a permuted qubit map or the test checks nothing"
    assert piece_ok
    np.testing.assert_allclose(full, ref, rtol=0, atol=atol)

    # sampler: identical to the reference sampler semantics run on the whole probability vector
    rdtype = np.float64 if dtype == "complex128" else np.float32
    probs = O.calculate_probabilities(full.astype(dtype), list(range(n)), n).astype(rdtype)
    for key, seed0, nshots in (("freqs", 7, 200000), ("nshots", 11, 150000)):
        np.random.seed(seed0)
        seed = int(np.random.randint(0, int(1e8), size=1, dtype=np.int64)[0])
        expect = np.zeros(1 << n, dtype=np.int64)
        O.measure_frequencies(expect, probs, nshots, n, seed, 4)
        got = np.zeros(1 << n, dtype=np.int64)
        for k, v in out[key].items():
            got[k] = v
        assert got.sum() == nshots
        if key == "freqs":
            np.testing.assert_array_equal(got, expect)      # bit-exact chain

    for tag, qubits, shot, normalize in [("c1", [0, n - 1], 2, True), ("c2", [1], 1, True),
                                         ("c3", [0, 2, 3], 5, False), ("c4", list(range(n)), 6, True)]:
        want = R.collapse(O, ref.copy(), qubits, shot, n, normalize)
        np.testing.assert_allclose(out[tag], want, rtol=0, atol=atol * 10, err_msg=tag)
        idx = np.arange(1 << n)                                  # the projected-out indices are exactly zero
        keep = np.ones(1 << n, dtype=bool)
        for j, qq in enumerate(sorted(qubits)):
            keep &= ((idx >> (n - 1 - qq)) & 1) == ((shot >> (len(qubits) - 1 - j)) & 1)
        assert not np.any(out[tag][~keep]) and not np.any(want[~keep]), tag

    init, final = out["init"]
    circuit = _circuit(n)
    want = R.reference_run(init, circuit.queue, n)
    np.testing.assert_allclose(final, want, rtol=0, atol=atol * 10)
