"""-m gpu tests of the distributed layer's device side on ONE rank: every gate becomes a
LocalGate in shard numbering, the local segment is compiled into a pass program
(`B200Backend.run_local_segment`) and must reproduce the einsum reference.  (The multi-rank NCCL
path is exercised by tools/dist_check.py under torchrun; its logs are under profiles/.)"""

import numpy as np
import pytest

from oracle import oracle as O_
from tests.gpu_utils import ATOL, backend
from tests.test_distributed_cpu import _reference_state, _test_circuits

pytestmark = pytest.mark.gpu


class _OneRank:
    rank, world = 0, 1

    def barrier(self):
        pass

    def all_reduce_sum(self, t):
        return t

    def all_gather(self, t):
        return [t]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("use_programs", [True, False])
def test_single_rank_distributed_state_matches_reference(dtype, use_programs):
    from qibojit_b200.distributed import DistributedState, LocalSegment

    b = backend()
    b.set_dtype(dtype)
    b.use_programs = use_programs
    try:
        n = 13
        for name, circuit in _test_circuits(n).items():
            ds = DistributedState(b, n, comm=_OneRank(), dtype=dtype)
            steps = ds.plan(circuit.queue)
            assert all(isinstance(s, LocalSegment) for s in steps)
            ds.run(steps)
            got = ds.to_numpy_full()
            ref = _reference_state(circuit, dtype)
            np.testing.assert_allclose(got, ref, rtol=0, atol=ATOL[dtype], err_msg=name)
            probs = b.to_numpy(ds.probabilities([0, n - 1, 2]))
            p = (np.abs(ref.astype(np.complex128)) ** 2).reshape((2,) * n)
            p = np.transpose(p.sum(axis=tuple(a for a in range(n) if a not in (0, n - 1, 2))), [0, 2, 1]).ravel()
            np.testing.assert_allclose(probs, p, rtol=1e-5 if dtype == "complex64" else 1e-12, atol=1e-7, err_msg=name)
            # the cached plan runs again from |0..0>
            ds.reset()
            ds.run(steps)
            np.testing.assert_array_equal(ds.to_numpy_full(), got, err_msg=name)
    finally:
        b.use_programs = True
        b.set_dtype("complex128")


def test_measurement_leg_of_the_bench():
    """bench.py's configs[4] leg (probabilities, 10^6 Metropolis shots, collapse) at a small size."""
    import bench
    from qibojit_b200 import circuits

    b = backend()
    b.set_dtype("complex128")
    n = 16
    state = b.execute_circuit(circuits.quantum_volume(n, depth=4))
    out = bench.measurement_leg(b, state, n, "complex128", nshots=10 ** 6)
    assert out["nshots"] == 10 ** 6 and out["distinct_outcomes"] > 1000
    assert abs(out["norm_after_collapse"] - 1.0) < 1e-9


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_single_rank_layout_and_measurement(dtype):
    """The sharded state's layout / measurement functions on one rank (world size 1, CUDA kernels):
    relabelled SWAPs leave a permuted qubit map, `to_tensor` undoes it ON THE DEVICE (a local segment
    of SWAP gates compiled into passes), `collapse` and `sample_frequencies` agree with the
    single-state backend (sampler: bit-exact).  The multi-rank versions run under torchrun
    (tools/dist_check.py) and on gloo (tests/test_distributed_measure_cpu.py)."""
    from qibojit_b200 import circuits, gates
    from qibojit_b200.circuit import Circuit
    from qibojit_b200.distributed import DistributedState, execute_distributed_circuit
    from tests import refdispatch as R

    b = backend()
    b.set_dtype(dtype)
    try:
        n = 13
        tol = ATOL[dtype]
        c = Circuit(n)
        c.add(circuits.qft(n).queue)
        c.add([gates.RY(0, 0.3), gates.CNOT(0, n - 1), gates.SWAP(1, n - 2), gates.H(n - 1), gates.CU1(0, 2, 0.4),
               gates.RX(1, 0.7), gates.SWAP(0, 2)])
        ref = _reference_state(c, dtype)
        ds = DistributedState(b, n, comm=_OneRank(), dtype=dtype)
        ds.execute(c.queue)
        assert ds.bit_of != [n - 1 - q for q in range(n)]
        full = ds.to_tensor()
        assert full.is_cuda and ds.bit_of == [n - 1 - q for q in range(n)]
        np.testing.assert_allclose(b.to_numpy(full), ref, rtol=0, atol=tol)
        np.random.seed(3)
        f1 = ds.sample_frequencies(200000)
        np.random.seed(3)
        f2 = b.sample_frequencies(b.calculate_probabilities(full.clone(), list(range(n)), n), 200000)
        assert sum(f1.values()) == 200000
        assert dict(f1) == dict(f2)                # the single-state sampler on the same amplitudes: same chains
        for qubits, shot in [([0, n - 1], 2), ([1], 1), ([0, 2, 3], 5)]:
            d2 = DistributedState(b, n, comm=_OneRank(), dtype=dtype)
            d2.execute(c.queue)
            d2.collapse(qubits, shot)
            want = R.collapse(O_, ref.copy(), qubits, shot, n, True)
            np.testing.assert_allclose(d2.to_numpy_full(), want, rtol=0, atol=tol * 10)
        init = R.random_state(n, dtype, 4)
        d3 = execute_distributed_circuit(b, c, initial_state=init, comm=_OneRank())
        np.testing.assert_allclose(b.to_numpy(d3.to_tensor()), R.reference_run(init, c.queue, n), rtol=0, atol=tol * 10)
        with pytest.raises(TypeError):
            execute_distributed_circuit(b, c, initial_state="zeros", comm=_OneRank())
    finally:
        b.set_dtype("complex128")


def test_multi_rank_parity_under_torchrun():
    """The whole of tools/dist_check.py on two ranks of this box (one process per GPU, NCCL + CUDA
    IPC): distributed parity on both exchange transports, layout / collapse / sampling on the sharded
    state, the pipelined out-of-place exchange bit for bit against the plain order, two devices
    driven from one process.  Needs two GPUs; a single-GPU box skips it."""
    import os
    import socket
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, QJ_NLOCAL="24")          # (small shards for the bandwidth section)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(root, "tools", "dist_check.py")],
                         capture_output=True, text=True, timeout=900, cwd=root, env=env)
    tail = "\n".join((out.stdout + out.stderr).splitlines()[-40:])
    assert out.returncode == 0, tail
    assert "rank 0: ALL OK" in out.stdout and "rank 1: ALL OK" in out.stdout, tail
    assert "FAIL" not in out.stdout, tail
