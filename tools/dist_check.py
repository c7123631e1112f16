#!/usr/bin/env python
"""Multi-GPU parity + timing check of the distributed layer (run under torchrun, NCCL).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

Every rank runs the test circuits of tests/test_distributed_cpu.py at a small size with both
exchange transports (peer memory over CUDA IPC, NCCL send/recv) and compares the gathered state with
a single-state reference; then the layout / measurement functions of the sharded state
(`to_tensor`, `collapse`, `sample_frequencies`, initial states) against the single-GPU backend
(sampler: bit-exact); then two devices driven from ONE process (the reference's joblib-thread
model, gpu.py:688-694); then the exchange bandwidth of both transports at a large shard size."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def say(rank, *a):
    if rank == 0:
        print(*a, flush=True)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from qibojit_b200 import circuits, gates
    from qibojit_b200.backends.b200 import B200Backend
    from qibojit_b200.circuit import Circuit
    from qibojit_b200.distributed import Comm, DistributedState, execute_distributed_circuit
    from tests import refdispatch as R
    from tests.test_distributed_cpu import _reference_state, _test_circuits

    b = B200Backend()
    ok = True
    # ---- circuits, both transports
    for transport in ("1", "0"):
        os.environ["QJ_PEER_EXCHANGE"] = transport
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
                say(rank, f"{'peer' if transport == '1' else 'nccl'} {dtype:10s} {name:18s} max|err| {err:.2e} "
                          f"exchanges {ds.stats['exchanges']} {'ok' if err < tol else 'FAIL'}")
                ok &= err < tol
    os.environ["QJ_PEER_EXCHANGE"] = "1"
    say(rank, "peer transport in use:", bool(getattr(b, "_peer_cache", None)), "broken:", getattr(b, "_peer_broken", False))

    # ---- layout + measurement on the sharded state vs the single-GPU backend
    for dtype in ("complex128", "complex64"):
        b.set_dtype(dtype)
        n = 14
        tol = 1e-5 if dtype == "complex64" else 1e-12
        c = Circuit(n)
        c.add(circuits.qft(n).queue)
        c.add([gates.RY(0, 0.3), gates.CNOT(0, n - 1), gates.SWAP(1, n - 2), gates.H(n - 1), gates.CU1(0, 2, 0.4),
               gates.RX(1, 0.7), gates.SWAP(0, 2)])
        ref = _reference_state(c, dtype)
        ds = DistributedState(b, n, comm=Comm(), dtype=dtype)
        ds.execute(c.queue)
        permuted = ds.bit_of != [n - 1 - q for q in range(n)]
        full = ds.to_tensor()
        err = float(np.abs(b.to_numpy(full) - ref).max())
        say(rank, f"{dtype} to_tensor (map was permuted: {permuted}) max|err| {err:.2e} {'ok' if err < tol else 'FAIL'}")
        ok &= err < tol and permuted
        # sampler: bit-exact with the single-GPU sampler on the same probabilities
        np.random.seed(7)
        f_dist = ds.sample_frequencies(300000)
        np.random.seed(7)
        single = b.cast(b.to_numpy(full), dtype=dtype, copy=True)
        f_one = b.sample_frequencies(b.calculate_probabilities(single, list(range(n)), n), 300000)
        same = dict(f_dist) == dict(f_one)
        say(rank, f"{dtype} sample_frequencies 3e5 shots, bit-exact with the single-GPU sampler: {same}")
        ok &= same
        for qubits, shot, normalize in [([0, n - 1], 2, True), ([1], 1, True), ([0, 2, 3], 5, False)]:
            d2 = DistributedState(b, n, comm=Comm(), dtype=dtype)
            d2.execute(c.queue)
            d2.collapse(qubits, shot, normalize=normalize)
            got = d2.to_numpy_full()
            one = b.cast(ref, dtype=dtype, copy=True)
            want = b.to_numpy(b.collapse_state(one, qubits, shot, n, normalize=normalize))
            err = float(np.abs(got - want).max())
            zeros = bool(np.array_equal(got == 0, want == 0))
            say(rank, f"{dtype} collapse {qubits}->{shot} normalize={normalize}: max|err| {err:.2e}, same zero pattern: {zeros} "
                      f"{'ok' if err < tol * 10 and zeros else 'FAIL'}")
            ok &= err < tol * 10 and zeros
        rng = np.random.default_rng(5)
        init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        init = (init / np.linalg.norm(init)).astype(dtype)
        d3 = execute_distributed_circuit(b, c, initial_state=init, comm=Comm())
        got = b.to_numpy(d3.to_tensor())
        want = R.reference_run(init, c.queue, n)
        err = float(np.abs(got - want).max())
        say(rank, f"{dtype} execute_distributed_circuit(initial_state) max|err| {err:.2e} {'ok' if err < tol * 10 else 'FAIL'}")
        ok &= err < tol * 10
        b.release_peer_mappings()
    b.set_dtype("complex128")

    # ---- pipelined exchange: the last pass of a segment overlapped with the out-of-place copy-engine
    # exchange, piece by piece, against the plain order (bit for bit) and the einsum reference
    for dtype in ("complex128", "complex64"):
        b.set_dtype(dtype)
        n = 20
        tol = 1e-5 if dtype == "complex64" else 1e-12
        from tests.circuits_random import random_circuit_gates

        cases = [("qft", circuits.qft(n)), ("supremacy", circuits.supremacy(n, depth=6)),
                 ("variational", circuits.variational(n))]
        # the last pass before the exchange works on the qubits that leave (all of them / one of them):
        # exchanged bits inside the pass's tile
        g = world.bit_length() - 1
        for part in (False, True):
            q = [gates.H(i) for i in range(g, n)]
            for i, t in enumerate(list(range(g, 2 * g))[: (1 if part else g)]):
                q += [gates.fSim(t, n - 1 - i, 0.3 + i, 0.7), gates.RY(t, 0.4)]
            q += [gates.H(i) for i in range(g)] + [gates.RY(i, 0.1 * i) for i in range(n) if not g <= i < 2 * g]
            cc = Circuit(n)
            cc.add(q)
            cases.append(("leaving-" + ("one" if part else "all"), cc))
        for seed in range(6):
            rc = Circuit(n)
            rc.add([gates.H(q) for q in range(n)] + random_circuit_gates(n, 150, 100 + seed))
            cases.append((f"random-{seed}", rc))
        for name, circuit in cases:
            outs = {}
            for mode in ("1", "0"):
                os.environ["QJ_OVERLAP_EXCHANGE"] = mode
                before = getattr(b, "overlapped_exchanges", 0)
                before_in = getattr(b, "overlapped_inside_tile", 0)
                ds = DistributedState(b, n, comm=Comm(), dtype=dtype)
                ds.execute(circuit.queue)
                outs[mode] = (ds.to_numpy_full(), getattr(b, "overlapped_exchanges", 0) - before,
                              getattr(b, "overlapped_inside_tile", 0) - before_in, ds.stats["exchanges"])
                b.release_peer_mappings()
            os.environ["QJ_OVERLAP_EXCHANGE"] = "1"
            same = bool(np.array_equal(outs["1"][0], outs["0"][0]))
            err = float(np.abs(outs["1"][0] - _reference_state(circuit, dtype)).max())
            # (the variational circuit's exchanges are not on the shard's top bits: it takes the plain order)
            good = same and err < tol and outs["0"][1] == 0 and (outs["1"][1] >= 1 or name not in ("qft", "supremacy"))
            if name.startswith("leaving"):
                good = good and outs["1"][2] >= 1
            say(rank, f"pipelined exchange {dtype:10s} {name:12s}: {outs['1'][3]} exchanges, pipelined {outs['1'][1]}x ({outs['1'][2]}x with exchanged bits inside the tile), identical to the plain order: {same}, "
                      f"max|err| vs reference {err:.2e} {'ok' if good else 'FAIL'}")
            ok &= good
    b.set_dtype("complex128")

    # ---- two devices from one process (rank 0 drives its own device and its neighbour's)
    if rank == 0 and torch.cuda.device_count() >= 2:
        n = 16
        circuit = circuits.qft(n)
        outs = []
        for dev in (0, 1, 0, 1):
            # (the process's current device stays 0 throughout: every library entry point switches to
            # its handle's device and back)
            bb = B200Backend(device=f"/GPU:{dev}")
            st = bb.execute_circuit(circuit)              # 32 KiB tiles: > 48 KiB dynamic shared memory per CTA
            st = bb.apply_gate(gates.H(3), st, n)
            st = bb.apply_gate(gates.H(3), st, n)
            outs.append(float(np.abs(bb.to_numpy(st) - 2.0 ** (-n / 2)).max()))
            assert st.device.index == dev
        cur = torch.cuda.current_device()
        good = max(outs) < 1e-12 and cur == 0
        print(f"two devices in one process: max|err| {max(outs):.2e}, current device left at {cur}: {'ok' if good else 'FAIL'}",
              flush=True)
        ok &= good
    dist.barrier()

    # ---- exchange bandwidth: single qubit (top bit: contiguous; middle bit) and the all-to-all over
    # the top log2(world) bits, both transports
    nlocal = int(os.environ.get("QJ_NLOCAL", "29"))
    g = world.bit_length() - 1
    for transport in ("1", "0"):
        os.environ["QJ_PEER_EXCHANGE"] = transport
        ds = DistributedState(b, nlocal + g, comm=Comm(), dtype="complex128")
        cases = [("1 qubit, top bit", lambda: b.shard_exchange(ds.shard, nlocal, nlocal - 1, rank ^ 1, rank & 1, ds.comm)),
                 ("1 qubit, bit 10", lambda: b.shard_exchange(ds.shard, nlocal, 10, rank ^ 1, rank & 1, ds.comm)),
                 (f"{g} qubits, top bits", lambda: b.shard_exchange_multi(
                     ds.shard, nlocal, list(range(nlocal - g, nlocal)), list(range(g)), rank, ds.comm))]
        for label, fn in cases:
            for rep in range(3):
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                moved = fn()
                e1.record(); torch.cuda.synchronize()
                if rank == 0 and rep == 2:
                    ms = e0.elapsed_time(e1)
                    print(f"{'peer' if transport == '1' else 'nccl'} exchange {label}: {moved / 2**30:.2f} GiB per rank and direction "
                          f"in {ms:.2f} ms = {moved / ms / 1e6:.1f} GB/s per direction", flush=True)
        dist.barrier()
        b.release_peer_mappings()
        dist.barrier()
        del ds
        torch.cuda.empty_cache()
    # ---- copy-engine pulls over the peer mapping (the transport of the pipelined exchange): every
    # rank pulls 2^(nlocal-1) amplitudes from its neighbour at the same time, in 1 / 2 / 4 pieces on
    # as many streams
    import ctypes

    from qibojit_b200 import _capi
    os.environ["QJ_PEER_EXCHANGE"] = "1"
    src = torch.zeros(1 << nlocal, dtype=torch.complex128, device=b.torch_device)
    dst = torch.empty(1 << (nlocal - 1), dtype=torch.complex128, device=b.torch_device)
    ptrs = b._peer_pointers(src, Comm())
    nbytes = dst.numel() * 16
    main_stream = torch.cuda.current_stream()
    for nstreams in (1, 2, 4):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        for rep in range(3):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            part = nbytes // nstreams
            for i, st in enumerate(streams):
                st.wait_stream(main_stream)
                _capi.check(b._lib.qj_set_stream(b._handle(), ctypes.c_void_p(st.cuda_stream)))
                _capi.check(b._lib.qj_copy_async(b._handle(), ctypes.c_void_p(dst.data_ptr() + i * part),
                                                 ctypes.c_void_p(ptrs[rank ^ 1] + i * part), part))
                main_stream.wait_stream(st)
            _capi.check(b._lib.qj_set_stream(b._handle(), ctypes.c_void_p(main_stream.cuda_stream)))
            e1.record(); torch.cuda.synchronize()
            if rank == 0 and rep == 2:
                ms = e0.elapsed_time(e1)
                print(f"copy-engine pull, {nstreams} stream(s): {nbytes / 2**30:.2f} GiB per rank in {ms:.2f} ms = "
                      f"{nbytes / ms / 1e6:.1f} GB/s per direction", flush=True)
    dist.barrier()
    b.release_peer_mappings()
    dist.barrier()
    del src, dst
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: {'ALL OK' if ok else 'FAILURES'}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
