#!/usr/bin/env python
"""Multi-GPU parity + timing check of the distributed layer (run under torchrun, NCCL).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

Every rank runs the test circuits of tests/test_distributed_cpu.py at a small size and
compares the gathered state with a single-state reference, then times half-shard exchanges at
a large shard size."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from qibojit_b200.backends.b200 import B200Backend
    from qibojit_b200.distributed import Comm, DistributedState
    from tests.test_distributed_cpu import _reference_state, _test_circuits

    b = B200Backend()
    ok = True
    for dtype in ("complex128", "complex64"):
        b.set_dtype(dtype)
        n = 12
        for name, circuit in _test_circuits(n).items():
            ds = DistributedState(b, n, comm=Comm(), dtype=dtype)
            ds.execute(circuit.queue)
            full = ds.to_numpy_full()
            ref = _reference_state(circuit, dtype)
            err = float(np.abs(full - ref).max())
            tol = 1e-5 if dtype == "complex64" else 1e-12
            if rank == 0:
                print(f"{dtype:10s} {name:18s} max|err| {err:.2e} exchanges {ds.stats['exchanges']} "
                      f"{'ok' if err < tol else 'FAIL'}", flush=True)
            ok &= err < tol
    # exchange bandwidth: top local bit (contiguous) and a middle bit (packed)
    b.set_dtype("complex128")
    nlocal = int(os.environ.get("QJ_NLOCAL", "29"))
    ds = DistributedState(b, nlocal + (world.bit_length() - 1), comm=Comm(), dtype="complex128")
    for lbit in (nlocal - 1, 10):
        peer = rank ^ 1
        for rep in range(3):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            moved = b.shard_exchange(ds.shard, nlocal, lbit, peer, rank & 1, ds.comm)
            e1.record(); torch.cuda.synchronize()
            if rank == 0 and rep:
                ms = e0.elapsed_time(e1)
                print(f"exchange lbit={lbit} {moved / 2**30:.1f} GiB per direction in {ms:.2f} ms = "
                      f"{moved / ms / 1e6:.1f} GB/s per direction", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
