"""A/B timing of the pipelined (out-of-place, copy-engine) exchange against the plain order on the
benchmark circuits: torchrun --nproc-per-node N tools/overlap_ab.py [workload:n:dtype ...]."""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from bench import build_circuit
    from qibojit_b200.backends.b200 import B200Backend
    from qibojit_b200.distributed import Comm, DistributedState

    b = B200Backend()
    cases = sys.argv[1:] or ["supremacy:34:complex64", "qft:33:complex128"]
    modes = os.environ.get("AB_MODES", "0,1:1,1:4,1:16").split(",")
    for case in cases:
        name, n, dtype = case.split(":")
        n = int(n)
        b.set_dtype(dtype)
        circuit = build_circuit(name, n)
        for mode in modes:
            on, _, slices = mode.partition(":")
            os.environ["QJ_OVERLAP_EXCHANGE"] = on
            os.environ["QJ_OVERLAP_SLICES"] = slices or "4"
            state = DistributedState(b, n, comm=Comm(), dtype=dtype)
            plan = state.plan(circuit.queue)
            before = getattr(b, "overlapped_exchanges", 0)
            for _ in range(2):
                state.reset(); state.run(plan)
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            steps = 3
            for _ in range(steps):
                state.reset(); state.run(plan)
            e1.record(); torch.cuda.synchronize(); dist.barrier()
            ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            from bench_distributed import _check_parity

            try:
                par = _check_parity(state, name, n, dtype, dist)
                norm = f"parity {par.get('max_abs_err')} ({'pinned' if par.get('pinned', True) else 'sum only'})"
            except AssertionError as exc:
                norm = f"PARITY FAILED: {exc}"
            if on == "1" and os.environ.get("QJ_OVERLAP_TRACE") == "1":
                state.reset(); state.run(plan)
                print(f"   rank {rank} trace {getattr(b, 'overlap_trace', None)}", flush=True)
            if rank == 0:
                print(f"{case} overlap={on} slices={slices or '-'}: {float(ms[0]):.1f} ms per circuit, "
                      f"pipelined exchanges {getattr(b, 'overlapped_exchanges', 0) - before}, {norm}", flush=True)
            state.shard = None
            state._spare = None
            del state
            dist.barrier()
            b.release_peer_mappings()
            dist.barrier()
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
