"""world_size-2/4 gloo tests of the distributed state layer on CPU ranks (host-side logic:
qubit map, rank predicates, diagonal restriction, swap scheduling, exchange pairing).  The
local kernels are the CPU oracle through tests/oracle_backend.py."""

import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_state(circuit, dtype):
    from qibojit_b200 import fusion
    from qibojit_b200.matrices import CustomMatrices
    from tests import refdispatch as R

    n = circuit.nqubits
    mats = CustomMatrices("complex128")
    st = np.zeros(1 << n, dtype=np.complex128)
    st[0] = 1
    for g in circuit.queue:
        if g.name == "fanout":
            for t in g.target_qubits:
                st = R.einsum_apply(st, mats.X, [t], [g.control_qubits[0]], n)
        else:
            st = R.einsum_apply(st, fusion.target_only_matrix(g, mats), list(g.target_qubits),
                                list(g.control_qubits), n)
    return st.astype(dtype)


def _test_circuits(n):
    from qibojit_b200 import circuits, gates
    from qibojit_b200.circuit import Circuit

    rng = np.random.default_rng(3)
    mixed = Circuit(n)
    mixed.add(gates.H(q) for q in range(n))
    mixed.add([gates.CNOT(0, n - 1), gates.CZ(1, 0), gates.CU1(0, 1, 0.3), gates.RZ(0, 0.7),
               gates.Z(0), gates.U1(1, 0.2), gates.RZZ(0, 1, 0.4), gates.RZZ(0, n - 1, 0.9),
               gates.CRZ(2, 0, 0.5), gates.CRZ(0, 1, 0.6), gates.TOFFOLI(0, 1, 2), gates.SWAP(0, n - 1),
               gates.fSim(0, 2, 0.3, 0.8), gates.CCZ(0, 1, 2), gates.Y(0), gates.CY(1, 0),
               gates.RX(1, 0.4), gates.SWAP(1, 2).controlled_by(0), gates.FanOut(0, 1, n - 1),
               gates.Unitary(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)), 0, n - 2),
               gates.Unitary(rng.standard_normal((8, 8)) + 0j, 1, n - 1, 0).controlled_by(2)])
    return {"qft": circuits.qft(n), "variational": circuits.variational(n),
            "variational_fused": circuits.variational(n).fuse(3), "supremacy": circuits.supremacy(n, depth=4),
            "qv": circuits.quantum_volume(n, depth=3), "mixed": mixed, "mixed_fused": mixed.fuse(3)}


def _worker(rank, world, port, n, dtype, q, via_planner=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from qibojit_b200.distributed import Comm, DistributedState
        from tests.oracle_backend import OracleBackend

        import torch

        O.set_threads(1)                     # `world` ranks share the host cores
        torch.set_num_threads(1)

        results = {}
        for name, circuit in _test_circuits(n).items():
            b = OracleBackend(dtype, via_planner=via_planner)
            ds = DistributedState(b, n, comm=Comm(), dtype=dtype)
            steps = ds.plan(circuit.queue)
            ds.run(steps)
            full = ds.to_numpy_full()
            probs = ds.probabilities([0, n - 1, 2]).numpy()
            results[name] = (full, probs, dict(ds.stats), ds.norm2())
            # a cached plan re-runs from |0..0> (bench loop), and program order gives the same state
            ds.reset()
            ds.run(steps)
            assert np.array_equal(ds.to_numpy_full(), full), name
            ds3 = DistributedState(b, n, comm=Comm(), dtype=dtype)
            ds3.run(ds3.plan(circuit.queue, batch_exchanges=False))
            np.testing.assert_allclose(ds3.to_numpy_full(), full, rtol=0,
                                       atol=1e-5 if dtype == "complex64" else 1e-12, err_msg=name)
            ds2 = DistributedState(b, n, comm=Comm(), dtype=dtype)
            ds2.run(ds2.plan(circuit.queue, reorder=False))
            np.testing.assert_allclose(ds2.to_numpy_full(), full, rtol=0,
                                       atol=1e-5 if dtype == "complex64" else 1e-12, err_msg=name)
            assert ds2.stats["exchanges"] >= ds.stats["exchanges"] // 2, name
        if rank == 0:
            q.put(results)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,via_planner,dtype", [
    (2, 6, False, "complex128"), (4, 7, False, "complex64"),
    (8, 8, False, "complex128"),          # three global qubits: 8-way all-to-all exchange
    (2, 8, True, "complex64"), (4, 9, True, "complex128")])
def test_distributed_state_matches_single_state(world, n, via_planner, dtype):
    """via_planner: each rank's local segments are lowered by planner.plan_queue (what the B200
    backend compiles into pass programs) and interpreted in numpy, instead of gate by gate."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world + 8 * via_planner
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, dtype, q, via_planner)) for r in range(world)]
    for p in procs:
        p.start()
    results = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    atol = 1e-5 if dtype == "complex64" else 1e-12
    for name, circuit in _test_circuits(n).items():
        full, probs, stats, norm2 = results[name]
        ref = _reference_state(circuit, dtype)
        np.testing.assert_allclose(full, ref, rtol=0, atol=atol, err_msg=name)
        p = (np.abs(ref.astype(np.complex128)) ** 2).reshape((2,) * n)
        unmeasured = tuple(a for a in range(n) if a not in (0, n - 1, 2))
        p = np.transpose(p.sum(axis=unmeasured), [0, 2, 1]).ravel()  # axes (0, 2, n-1) -> (0, n-1, 2)
        np.testing.assert_allclose(probs, p, rtol=0, atol=1e-5 if dtype == "complex64" else 1e-12, err_msg=name)
        assert abs(norm2 - float(np.vdot(ref, ref).real)) < 1e-4
    # the QFT's controlled phases and final swaps must not trigger exchanges beyond the H gates
    assert results["qft"][2]["exchanges"] == 1       # one (multi-)exchange whatever the world size
    assert results["qft"][2]["relabelled_swaps"] == n // 2


def test_lookahead_prefers_far_victims():
    """Single-rank logic check of the swap victim choice (no process group needed)."""
    from qibojit_b200.distributed import DistributedState

    class FakeComm:
        rank, world = 0, 1

        def barrier(self):
            pass

    from tests.oracle_backend import OracleBackend

    ds = DistributedState(OracleBackend(), 5, comm=FakeComm())
    victim = ds._choose_victim({0}, {1: 3, 2: 50, 3: 7})
    assert victim == 4  # never used again -> farthest
    victim = ds._choose_victim({0, 4}, {1: 3, 2: 50, 3: 7})
    assert victim == 2


def test_dag_schedule_needs_few_exchanges():
    """Exchange counts of the benchmark circuits (planning only, no amplitudes)."""
    from qibojit_b200 import circuits
    from qibojit_b200.distributed import DistributedState, Exchange
    from tests.oracle_backend import OracleBackend

    class FakeComm:
        def __init__(self, rank, world):
            self.rank, self.world = rank, world

    class NoAlloc(OracleBackend):
        def shard_zeros(self, nlocal, dtype, one_at_zero=False):
            return None

    def count(circuit, n, world, dtype, **kw):
        """(number of exchange steps, shards moved) -- identical on every rank."""
        from qibojit_b200.distributed import MultiExchange

        per_rank = []
        for rank in (0, world - 1):
            ds = DistributedState(NoAlloc(dtype), n, comm=FakeComm(rank, world), dtype=dtype)
            steps = ds.plan(circuit.queue, **kw)
            sig = []
            for s in steps:
                if isinstance(s, Exchange):
                    sig.append(((s.rank_bit,), (s.local_bit,)))
                elif isinstance(s, MultiExchange):
                    sig.append((tuple(s.rank_bits), tuple(s.local_bits)))
            per_rank.append(sig)
        assert per_rank[0] == per_rank[1]      # every rank plans the same exchanges
        return len(per_rank[0]), sum(1 - 2.0 ** -len(r) for r, _ in per_rank[0])

    assert count(circuits.variational(30), 30, 2, "complex128") == (1, 0.5)
    assert count(circuits.variational(30), 30, 2, "complex128", reorder=False)[0] >= 5
    assert count(circuits.qft(34), 34, 2, "complex128") == (1, 0.5)
    # three global qubits: one all-to-all (7/8 of a shard) instead of three swaps (3/2)
    assert count(circuits.supremacy(36), 36, 8, "complex64") == (1, 0.875)
    assert count(circuits.qft(36), 36, 8, "complex64") == (1, 0.875)
    assert count(circuits.supremacy(36), 36, 8, "complex64", batch_exchanges=False) == (3, 1.5)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_plans_fuzz_in_one_process(world):
    """Random circuits (controls / diagonals / SWAPs on global qubits included) through every
    rank's plan, run in lockstep in one process (tests/virtual_ranks.py), in the three schedules."""
    from tests import refdispatch as R
    from tests.circuits_random import random_circuit_gates
    from tests.virtual_ranks import run_virtual

    rng = np.random.default_rng(world)
    for case in range(12):
        n = int(rng.integers(world.bit_length() + 3, 9))
        dtype = ["complex128", "complex64"][case % 2]
        glist = [g for g in random_circuit_gates(n, int(rng.integers(10, 60)), int(rng.integers(0, 1 << 30)))]
        st = np.zeros(1 << n, dtype=np.complex128)
        st[0] = 1
        ref = R.reference_run(st, glist, n)
        for kw in ({}, {"batch_exchanges": False}, {"reorder": False}):
            got, _ = run_virtual(glist, n, world, dtype, **kw)
            np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12 if dtype == "complex128" else 2e-5,
                                       err_msg=f"case {case} n={n} {dtype} {kw}")


def test_tile_split_over_exchanged_bits():
    """`B200Backend._tile_split` (host logic of the pipelined exchange): the exchanged top bits that
    lie outside a launch's tile enumerate contiguous tile ranges; checked against the kernel's tile
    numbering (index bits outside the tile, least significant first) by brute force."""
    import numpy as np

    from qibojit_b200.backends.b200 import B200Backend

    split = B200Backend._tile_split
    nlocal = 14
    rng = np.random.default_rng(3)
    idx = np.arange(1 << nlocal)
    for trial in range(60):
        r = int(rng.integers(2, 6))
        nh = int(rng.integers(0, 4))
        hibits = sorted(int(b) for b in rng.choice(np.arange(r, nlocal), size=nh, replace=False))
        T = r + nh
        geom = {"T": T, "r": r, "ntiles": 1 << (nlocal - T), "hibits": hibits}
        outside = [b for b in range(r, nlocal) if b not in hibits]
        tile = np.zeros_like(idx)
        for j, b in enumerate(outside):
            tile |= ((idx >> b) & 1) << j
        for ntop in (1, 2, 3):
            got = split(geom, nlocal, ntop)
            assert got is not None
            free, per = got
            assert free == [i for i in range(ntop) if (nlocal - ntop + i) not in hibits]
            assert per << len(free) == geom["ntiles"]
            value = np.zeros_like(idx)
            for j, i in enumerate(free):
                value |= ((idx >> (nlocal - ntop + i)) & 1) << j
            assert np.array_equal(tile // per, value)      # tile range v <-> free bits spell v
    # the contiguous part of the tile reaches into the exchanged bits: no split
    assert split({"T": 12, "r": 12, "ntiles": 4, "hibits": []}, 14, 3) is None
    assert split({"T": 12, "r": 11, "ntiles": 4, "hibits": [12]}, 14, 3) == ([0, 2], 1)
