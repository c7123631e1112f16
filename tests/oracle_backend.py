"""TEST INFRASTRUCTURE: a numpy/oracle stand-in with the backend surface the distributed layer
uses, so `qibojit_b200.distributed` can be exercised on CPU ranks (gloo, world_size 2+).
Kernels = the CPU oracle; shard exchange = torch.distributed send/recv on CPU tensors."""

import numpy as np
import torch

from oracle import oracle as O
from qibojit_b200.matrices import CustomMatrices
from tests import refdispatch as R


class OracleBackend:
    def __init__(self, dtype="complex128", via_planner=False):
        self.dtype = dtype
        self.via_planner = via_planner   # lower local segments through planner.plan_queue
        self.custom_matrices = CustomMatrices(dtype)
        self.engine = torch

    # --- same helpers as B200Backend
    def _as_custom_matrix(self, gate):
        from qibojit_b200 import fusion

        name = gate.__class__.__name__
        if name == "FusedGate":
            return fusion.fused_matrix(gate, self.custom_matrices)
        if name == "FanOut":
            return None
        return gate.target_matrix(self.custom_matrices)

    def to_numpy(self, x):
        return x.numpy() if hasattr(x, "numpy") else np.asarray(x)

    def _one_qubit_base(self, state, nqubits, target, kernel, gate, qubits):
        R.one_qubit_base(O, state.numpy(), nqubits, target, kernel, gate, qubits)
        return state

    def _two_qubit_base(self, state, nqubits, t1, t2, kernel, gate, qubits):
        R.two_qubit_base(O, state.numpy(), nqubits, t1, t2, kernel, gate, qubits)
        return state

    def _multi_qubit_base(self, state, nqubits, targets, gate, qubits):
        R.multi_qubit_base(O, state.numpy(), nqubits, list(targets), gate, qubits)
        return state

    def zero_state(self, nqubits, dtype=None):
        st = np.empty(1 << nqubits, dtype=dtype or self.dtype)
        return torch.from_numpy(O.initial_state_vector(st))

    def calculate_probabilities(self, state, qubits, nqubits):
        return torch.from_numpy(O.calculate_probabilities(state.numpy(), list(qubits), nqubits))

    def calculate_norm(self, state):
        return float(np.linalg.norm(state.numpy()))

    def collapse_state(self, state, qubits, shot, nqubits, normalize=True, density_matrix=False):
        R.collapse(O, state.numpy(), sorted(qubits), shot, nqubits, normalize)
        return state

    def sample_frequencies(self, probabilities, nshots):
        """Same seed draw and sampler as cpu.py:383-394 above the Metropolis threshold."""
        from collections import Counter

        seed = int(np.random.randint(0, int(1e8), size=1, dtype=np.int64)[0])
        probs = probabilities.numpy()
        nq = int(probs.size).bit_length() - 1
        freqs = np.zeros(probs.size, dtype=np.int64)
        O.measure_frequencies(freqs, probs, int(nshots), nq, seed, 4)
        nz = np.nonzero(freqs)[0]
        return Counter(dict(zip(nz.tolist(), freqs[nz].tolist())))

    # --- shard primitives
    def shard_from(self, piece, dtype):
        arr = piece.numpy() if hasattr(piece, "numpy") else np.asarray(piece)
        return torch.from_numpy(np.array(arr, dtype=dtype, copy=True).reshape(-1))

    def shard_zeros(self, nlocal, dtype, one_at_zero=False):
        if one_at_zero:
            return self.zero_state(nlocal, dtype)
        return torch.zeros(1 << nlocal, dtype=getattr(torch, str(dtype)))

    def shard_reset(self, shard, nlocal, one_at_zero=False):
        shard.zero_()
        if one_at_zero:
            shard[0] = 1
        return shard

    def run_local_segment(self, shard, nlocal, segment):
        if self.via_planner and nlocal >= 6:
            from qibojit_b200 import planner
            from tests import plan_interp

            plan = planner.plan_queue(segment.gates, nlocal, self.custom_matrices, 6, 3, dtype=self.dtype)
            st = plan_interp.run_plan(shard.numpy().astype(np.complex128), plan, nlocal,
                                      lambda st, g: g.apply(self, torch.from_numpy(st), nlocal).numpy())
            shard.copy_(torch.from_numpy(st.astype(self.dtype)))
            return shard
        for gate in segment.gates:
            shard = gate.apply(self, shard, nlocal)
        return shard

    def shard_scale(self, shard, nlocal, phase):
        shard.mul_(complex(phase))
        return shard

    def shard_exchange(self, shard, nlocal, lbit, peer, is_upper, comm, chunk_bytes=1 << 29):
        dist = comm.dist
        idx = np.arange(1 << nlocal)
        sel = torch.from_numpy(idx[((idx >> lbit) & 1) == (0 if is_upper else 1)])
        send = torch.view_as_real(shard[sel].contiguous())
        recv = torch.empty_like(send)
        if comm.rank < peer:
            dist.send(send, peer, group=comm.group)
            dist.recv(recv, peer, group=comm.group)
        else:
            dist.recv(recv, peer, group=comm.group)
            dist.send(send, peer, group=comm.group)
        shard[sel] = torch.view_as_complex(recv)
        return send.numel() * send.element_size()

    def shard_exchange_multi(self, shard, nlocal, lbits, rank_bits, rank, comm, chunk_bytes=1 << 29):
        """Pairwise sub-block swaps in XOR-distance order (both partners meet at the same step)."""
        dist = comm.dist
        k = len(lbits)
        mine = sum(((rank >> j) & 1) << i for i, j in enumerate(rank_bits))
        idx = np.arange(1 << nlocal)
        field = np.zeros_like(idx)
        for i, l in enumerate(lbits):
            field |= ((idx >> l) & 1) << i
        moved = 0
        for d in range(1, 1 << k):
            a = mine ^ d
            peer = rank
            for i, j in enumerate(rank_bits):
                peer = (peer & ~(1 << j)) | (((a >> i) & 1) << j)
            sel = torch.from_numpy(idx[field == a])
            send = torch.view_as_real(shard[sel].contiguous())
            recv = torch.empty_like(send)
            if rank < peer:
                dist.send(send, peer, group=comm.group)
                dist.recv(recv, peer, group=comm.group)
            else:
                dist.recv(recv, peer, group=comm.group)
                dist.send(send, peer, group=comm.group)
            shard[sel] = torch.view_as_complex(recv)
            moved += send.numel() * send.element_size()
        return moved
