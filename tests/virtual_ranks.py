"""TEST INFRASTRUCTURE: all ranks of a DistributedState in ONE process (no process group): every
rank plans the gate list for itself, the steps run in lockstep and exchanges are carried out
directly between the numpy shards.  Fast enough to fuzz the distributed planner."""

import numpy as np
import torch

from qibojit_b200.distributed import DistributedState, Exchange, LocalSegment, MultiExchange
from tests.oracle_backend import OracleBackend


class _Comm:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def barrier(self):
        pass


def run_virtual(queue, nqubits, world, dtype="complex128", **plan_kw):
    """-> (full state vector in logical qubit order, plan of rank 0)."""
    states = [DistributedState(OracleBackend(dtype), nqubits, comm=_Comm(r, world), dtype=dtype) for r in range(world)]
    plans = [s.plan(queue, **plan_kw) for s in states]
    nlocal = states[0].nlocal
    idx = np.arange(1 << nlocal)
    # a rank whose share of the gates between two exchanges is empty has no LocalSegment there:
    # align the plans on their exchanges, which must be identical on every rank
    def signature(q):
        return ([q.rank_bit], [q.local_bit]) if isinstance(q, Exchange) else (q.rank_bits, q.local_bits)

    exchanges = [[signature(q) for q in p if not isinstance(q, LocalSegment)] for p in plans]
    assert all(e == exchanges[0] for e in exchanges), "ranks disagree on the exchanges"
    cursors = [0] * world
    for k in range(len(exchanges[0]) + 1):
        for r, (s, p) in enumerate(zip(states, plans)):
            while cursors[r] < len(p) and isinstance(p[cursors[r]], LocalSegment):
                s.shard = s.backend.run_local_segment(s.shard, nlocal, p[cursors[r]])
                cursors[r] += 1
        if k == len(exchanges[0]):
            break
        rank_bits, local_bits = exchanges[0][k]
        for r in range(world):
            assert isinstance(plans[r][cursors[r]], (Exchange, MultiExchange))
            cursors[r] += 1
        old = [s.shard.numpy().copy() for s in states]
        field = np.zeros_like(idx)
        for i, l in enumerate(local_bits):
            field |= ((idx >> l) & 1) << i
        for r in range(world):
            mine = sum(((r >> j) & 1) << i for i, j in enumerate(rank_bits))
            new = old[r].copy()
            for a in range(1 << len(rank_bits)):
                if a == mine:
                    continue
                peer = r
                for i, j in enumerate(rank_bits):
                    peer = (peer & ~(1 << j)) | (((a >> i) & 1) << j)
                # this rank's sub-block `a` is replaced by the peer's sub-block `mine`
                new[field == a] = old[peer][field == mine]
            states[r].shard = torch.from_numpy(new)
    assert all(c == len(p) for c, p in zip(cursors, plans))
    for s, p in zip(states, plans):
        s.bit_of = list(p.final_map)
    assert len({tuple(s.bit_of) for s in states}) == 1, "ranks disagree on the final qubit map"
    phys = np.concatenate([s.shard.numpy() for s in states])
    n = nqubits
    axes = [n - 1 - states[0].bit_of[q] for q in range(n)]
    return np.transpose(phys.reshape((2,) * n), axes).reshape(-1), plans[0]
