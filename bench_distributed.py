"""N > 1 leg of bench.py: the workload circuit with its state sharded over the ranks (strong
scaling: the circuit and its 2^n amplitudes are fixed, each rank holds 2^n / N).

Primary workload: the supremacy-style circuit on 34 qubits in complex64 (137 GB: it fits one GPU,
so the N = 1 point of the series exists -- `secondary.supremacy` of the `--gpus 1` line).  Secondary
records of the same JSON line: QFT-33 complex128 sharded (BASELINE configs[2] on N GPUs) and, on 8
ranks, supremacy-36 complex64 (configs[3], 550 GB of state).  Every record carries its parity check:
QFT against the closed form on every shard, supremacy against the committed 4-qubit marginal of the
1-GPU run (tests/golden/marginals.json)."""

import json
import time

import numpy as np


class TimedBackend:
    """Proxy that brackets every kernel launch / exchange of the distributed layer with CUDA
    events (recorded on the launch stream) and forwards everything else."""

    def __init__(self, backend, nlocal, amp_bytes):
        self._b = backend
        self._nbytes = amp_bytes << nlocal
        self.records = []
        self.enabled = False

    def __getattr__(self, name):
        return getattr(self._b, name)

    def _timed(self, kind, alg, fn, *a, **k):
        if not self.enabled:
            return fn(*a, **k)
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        self.records.append((kind, alg, e0, e1))
        return out

    def run_local_segment(self, shard, nlocal, segment):
        if not self.enabled or segment.compiled is None:
            return self._b.run_local_segment(shard, nlocal, segment)
        return segment.compiled.run_timed(
            shard, lambda kind, frac, fn: self._timed(kind, 2.0 * self._nbytes * frac, fn))

    def run_segment_then_exchange(self, shard, nlocal, segment, lbits, rank_bits, rank, comm, chunk_bytes=1 << 29,
                                  spare=None):
        """Timed steps: the backend's pipelined path (last pass overlapped with the exchange).  The
        breakdown step (`enabled`): segment, then exchange, each launch bracketed by events."""
        if not self.enabled:
            return self._b.run_segment_then_exchange(shard, nlocal, segment, lbits, rank_bits, rank, comm,
                                                     chunk_bytes, spare=spare)
        shard = self.run_local_segment(shard, nlocal, segment)
        if len(lbits) == 1:
            peer = rank ^ (1 << rank_bits[0])
            moved = self.shard_exchange(shard, nlocal, lbits[0], peer, (rank >> rank_bits[0]) & 1, comm, chunk_bytes)
        else:
            moved = self.shard_exchange_multi(shard, nlocal, lbits, rank_bits, rank, comm, chunk_bytes)
        return shard, moved

    def shard_exchange_multi(self, shard, nlocal, lbits, rank_bits, rank, comm, chunk_bytes=1 << 29):
        return self._timed("exchange", self._nbytes * (1.0 - 2.0 ** -len(lbits)), self._b.shard_exchange_multi,
                           shard, nlocal, lbits, rank_bits, rank, comm, chunk_bytes)

    def shard_exchange(self, shard, nlocal, lbit, peer, is_upper, comm, chunk_bytes=1 << 29):
        return self._timed("exchange", self._nbytes / 2, self._b.shard_exchange, shard, nlocal, lbit,
                           peer, is_upper, comm, chunk_bytes)


def _check_parity(state, workload, nqubits, dtype, dist):
    """Parity of the sharded final state at benchmark size; raises on a mismatch above tolerance."""
    import torch

    from bench import MARGINAL_QUBITS, TOL, marginal_fixture

    b = state.backend
    tol = TOL[dtype]
    out = {"tolerance": tol}
    if workload == "qft":
        err = torch.tensor([b.max_deviation(state.shard, 2.0 ** (-nqubits / 2))], device="cuda", dtype=torch.float64)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        out.update(check="max |amp - 2^(-n/2)| over every shard (closed form of QFT|0..0>), on the devices",
                   max_abs_err=float(err[0]))
        if not out["max_abs_err"] <= tol:
            raise AssertionError(f"parity: sharded qft-{nqubits} deviates by {out['max_abs_err']:.3e} > {tol}")
        return out
    marg = state.probabilities(MARGINAL_QUBITS).double().cpu().numpy()
    out["marginal_sum"] = float(marg.sum())
    fx = marginal_fixture(workload, nqubits, dtype)
    if fx is None:
        out.update(check="4-qubit marginal sums to 1 (no reference marginal at this size)", pinned=False)
        if abs(out["marginal_sum"] - 1.0) > (1e-9 if dtype == "complex128" else 1e-4):
            raise AssertionError(f"parity: marginal of sharded {workload}-{nqubits} sums to {out['marginal_sum']}")
        return out
    ref, src = fx
    err = float(np.abs(marg - ref).max())
    out.update(check=f"4-qubit marginal (qubits {MARGINAL_QUBITS}) vs {src}", max_abs_err=err, pinned=True)
    if not err <= tol:
        raise AssertionError(f"parity: marginal of sharded {workload}-{nqubits} {dtype} differs by {err:.3e} > {tol}")
    return out


def measurement_leg_sharded(state, nqubits, dtype, nshots=10 ** 6):
    """BASELINE.json configs[4] on a sharded state: layout normalisation + full-register
    probabilities (all-gathered), 10^6-shot sampling with the reference's Metropolis sampler
    semantics (ops.py:86-108; every rank runs the same chains on the gathered vector) and one
    collapse on three qubits (local zeroing / rank predicate, all-reduced norm, rescale).  Each
    part is timed with CUDA events on the launch stream, max over ranks."""
    import torch
    import torch.distributed as dist

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return out, float(ms[0])

    np.random.seed(123)
    probs, ms_probs = timed(state.full_probabilities)
    freqs, ms_sample = timed(lambda: state.backend.sample_frequencies(probs, nshots))
    assert sum(freqs.values()) == nshots, sum(freqs.values())
    del probs
    torch.cuda.empty_cache()
    shot = max(freqs, key=freqs.get)
    qubits = [1, nqubits // 2, nqubits - 2]
    outcome = sum(((shot >> (nqubits - 1 - q)) & 1) << (len(qubits) - 1 - i) for i, q in enumerate(qubits))
    _, ms_collapse = timed(lambda: state.collapse(qubits, outcome))
    norm2 = state.norm2()
    assert abs(norm2 - 1.0) < (1e-9 if dtype == "complex128" else 1e-4), norm2
    return {"nshots": nshots, "distinct_outcomes": len(freqs),
            "layout_and_probabilities_ms": ms_probs, "sample_frequencies_ms": ms_sample,
            "shots_per_second": nshots / (ms_sample * 1e-3), "collapse_qubits": qubits, "collapse_ms": ms_collapse,
            "norm2_after_collapse": norm2,
            "note": "probabilities: device-side layout normalisation (exchange + local SWAP passes), |amp|^2 and an "
                    "all-gather of 2^n reals to every rank; the sampler then runs on every rank (same seed, same result)"}


def time_sharded(backend, workload, nqubits, dtype, steps, warmup, world, rank, measure=False):
    """Plan once, then time `steps` executions of the sharded circuit (device events, max over
    ranks).  Returns the record (same fields as bench.time_program) on every rank."""
    import torch
    import torch.distributed as dist

    from bench import ClockSampler, build_circuit, measured_peak_gbs
    from qibojit_b200.distributed import Comm, DistributedState, LocalSegment

    amp = 16 if dtype == "complex128" else 8
    backend.set_dtype(dtype)
    circuit = build_circuit(workload, nqubits)
    ngates = circuit.ngates
    comm = Comm()
    nlocal = nqubits - (world.bit_length() - 1)
    tb = TimedBackend(backend, nlocal, amp)
    state = DistributedState(tb, nqubits, comm=comm, dtype=dtype)
    # plan once: exchanges by look-ahead, this rank's local gates between them compiled into
    # multi-gate pass programs on first use (the warm-up steps)
    t0 = time.perf_counter()
    steps_plan = state.plan(circuit.queue)
    plan_ms = 1e3 * (time.perf_counter() - t0)

    def step():
        state.reset()          # re-prepare |0..0> in place
        state.run(steps_plan)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    launches0 = backend.launch_count()
    pipelined0 = getattr(backend, "overlapped_exchanges", 0)
    for k in state.stats:
        state.stats[k] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(int(torch.cuda.current_device())) as clocks:
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms[0]) / steps
    launches = backend.launch_count() - launches0
    pipelined = getattr(backend, "overlapped_exchanges", 0) - pipelined0
    stats = dict(state.stats)
    # per-kernel breakdown from two extra steps with every launch and exchange bracketed by events
    # (not pipelined: the timed steps above overlap the last pass of a segment with the exchange)
    tb.enabled = True
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    tb.enabled = False
    nbreak = 2

    per_kind = {}
    for kind, alg, a, b in tb.records:
        d = per_kind.setdefault(kind, {"ms": 0.0, "bytes": 0.0, "n": 0})
        d["ms"] += a.elapsed_time(b)
        d["bytes"] += alg
        d["n"] += 1
    local = {k: v for k, v in per_kind.items() if k != "exchange"}
    dom = max(local, key=lambda k: local[k]["ms"])
    peak, peak_src = measured_peak_gbs()
    achieved = local[dom]["bytes"] / (local[dom]["ms"] * 1e-3) / 1e9
    breakdown = {k: {"launches_per_step": v["n"] / nbreak, "avg_ms": v["ms"] / v["n"],
                     "ms_per_step": v["ms"] / nbreak, "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9}
                 for k, v in per_kind.items()}
    serial_ms = sum(v["ms"] for v in per_kind.values()) / nbreak
    parity = _check_parity(state, workload, nqubits, dtype, dist)
    measurement = None
    if measure:
        state.backend = backend          # (untimed proxy off: the leg brackets its own parts)
        measurement = measurement_leg_sharded(state, nqubits, dtype)
    h2d = sum(st.compiled.upload_bytes for st in steps_plan if isinstance(st, LocalSegment) and st.compiled is not None)
    record = {
        "workload": f"{workload}-{nqubits}-{dtype}", "circuit_gates": ngates, "n_gpus": world,
        "ms_per_step": ms_per_step, "value": ngates / (ms_per_step * 1e-3),
        "plan_ms": plan_ms, "local_segments": sum(isinstance(st, LocalSegment) for st in steps_plan),
        "shard_bytes": amp << nlocal, "exchanges_per_step": stats["exchanges"] / steps,
        "exchange_bytes_per_rank_per_step": stats["exchange_bytes"] / steps,
        "exchange_transport": ("peer memory (CUDA IPC over NVLink): copy-engine pulls into a second shard buffer, "
                               "pipelined under the last pass of the segment" if pipelined else
                               "peer memory (CUDA IPC over NVLink, in-place swap kernel)")
                              if getattr(backend, "_peer_cache", None) else "NCCL send/recv through staging buffers",
        "pipelined_exchanges_per_step": pipelined / steps,
        "gpu_launches": int(launches), "program_upload_bytes": int(h2d),
        "roofline": {"bound": "hbm", "kernel": "k_pass (multi-gate tile pass, 2*shard bytes per launch)" if dom == "pass" else dom,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "per_kernel": breakdown,
                     "per_kernel_note": "from two extra, un-pipelined steps; the timed steps overlap the last pass of a "
                                        "segment with the exchange (sum of the parts %.1f ms vs ms_per_step)" % serial_ms,
                     "exchange_note": "exchange GB/s = bytes sent per rank / time; NVLink reference 770 GB/s per direction"},
        "parity": parity, "clocks": clocks.summary(),
    }
    if measurement is not None:
        record["measurement"] = measurement
    torch.cuda.synchronize()
    dist.barrier()
    backend.release_peer_mappings()        # before any rank frees its shard
    dist.barrier()
    state.shard = None
    del state
    torch.cuda.empty_cache()
    return record, circuit


def time_sharded_e2e(backend, circuit, nqubits, dtype, reps):
    """End to end: gate objects in, marginal probabilities out; plans, compiles and uploads inside
    the timed region."""
    import torch
    import torch.distributed as dist

    from bench import MARGINAL_QUBITS
    from qibojit_b200.distributed import Comm, DistributedState

    comm = Comm()
    times, host = [], None
    # two untimed executions first: the shard and the spare buffer of the out-of-place exchange swap
    # roles from one execution to the next, and a peer's allocation is mapped (CUDA IPC, ~3 ms per
    # GiB, once per allocation and process) the first time it is the source of a pull
    warm = 2
    for i in range(warm + reps):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ds = DistributedState(backend, nqubits, comm=comm, dtype=dtype)
        ds.execute(circuit.queue)
        host = ds.probabilities(MARGINAL_QUBITS).cpu().numpy()
        torch.cuda.synchronize()
        dist.barrier()
        if i >= warm:
            times.append(time.perf_counter() - t0)
        ds.shard = None     # back to torch's caching allocator: the next step reuses the block
        del ds
    dist.barrier()
    backend.release_peer_mappings()
    dist.barrier()
    t = torch.tensor([float(np.mean(times))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert abs(host.sum() - 1.0) < (1e-6 if dtype == "complex128" else 1e-3), host.sum()
    return {"value": circuit.ngates / float(t[0]), "unit": "gates/s", "ms_per_step": 1e3 * float(t[0]),
            "d2h_bytes_per_step": int(host.nbytes),
            "includes": "planning, program encode + upload on every rank, state preparation, all passes and "
                        "exchanges, marginal all-reduce + read-back (the final state stays sharded on the devices)"}


def run_distributed(args, backend, world, rank):
    import torch
    import torch.distributed as dist

    from bench import DEFAULTS

    workload = args.workload or "supremacy"
    cfg = DEFAULTS[workload]
    nqubits = args.nqubits or cfg["nqubits"]
    dtype = args.dtype or cfg["dtype"]

    # ---- the N = 1 point of the same circuit (strong scaling): rank 0's GPU alone, unsharded
    # multi-gate passes, before any shard is allocated; the other ranks wait
    n1 = None
    amp = 16 if dtype == "complex128" else 8
    if not getattr(args, "no_n1", False) and (amp << nqubits) <= 150e9:
        if rank == 0:
            try:
                from bench import check_parity, time_program

                rec, st, _ = time_program(backend, workload, nqubits, dtype, min(args.steps, 2), 1)
                par = check_parity(backend, st, workload, nqubits, dtype)
                del st
                n1 = {"value": rec["value"], "unit": "gates/s", "ms_per_step": rec["ms_per_step"], "passes": rec["passes"],
                      "parity_max_abs_err": par.get("max_abs_err"),
                      "how": "the same circuit on rank 0's GPU alone (unsharded multi-gate passes), timed in this "
                             "process before the sharded run; also `secondary` of the --gpus 1 line"}
            except AssertionError:
                raise
            except Exception as exc:      # (e.g. the device is shared and short of memory)
                n1 = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
        dist.barrier()

    primary, circuit = time_sharded(backend, workload, nqubits, dtype, args.steps, args.warmup, world, rank,
                                    measure=(workload == "qv" or args.measure))
    e2e = time_sharded_e2e(backend, circuit, nqubits, dtype, min(args.steps, 2))
    e2e["h2d_bytes_per_step"] = primary["program_upload_bytes"]
    torch.cuda.empty_cache()

    secondary = {}
    if not args.no_secondary and not args.workload:
        extra = [("qft", 33, "complex128")]
        if world == 8:
            extra.append(("supremacy", 36, "complex64"))
        for name, n, dt in extra:
            try:
                rec, _ = time_sharded(backend, name, n, dt, min(args.steps, 3), 3, world, rank)
            except AssertionError:
                raise
            except Exception as exc:
                rec = {"error": f"{type(exc).__name__}: {exc}"}
            secondary[f"{name}-{n}"] = rec
            torch.cuda.empty_cache()

    line = None
    if rank == 0:
        line = {
            "metric": "gates_per_second", "value": primary["value"], "unit": "gates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": primary["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if dtype == "complex128" else "f32", "data": "synthetic",
            "config": {"workload": primary["workload"], "circuit_gates": primary["circuit_gates"],
                       "execution": "per-rank multi-gate tile passes between exchanges", "plan_ms": primary["plan_ms"],
                       "local_segments": primary["local_segments"], "shard_bytes": primary["shard_bytes"],
                       "exchanges_per_step": primary["exchanges_per_step"],
                       "exchange_bytes_per_rank_per_step": primary["exchange_bytes_per_rank_per_step"],
                       "exchange_transport": primary["exchange_transport"],
                       "pipelined_exchanges_per_step": primary.get("pipelined_exchanges_per_step"),
                       "parallelism": f"state sharded over {world} ranks on the top {world.bit_length() - 1} qubits; "
                                      "qubit exchanges over NVLink: all-to-all of (2^k-1)/2^k of a shard for k qubits "
                                      "(half a shard for one), out of place by the copy engines under the last pass "
                                      "when a second shard buffer fits, else by the in-place swap kernel",
                       "strong_scaling_n1": n1 if n1 is not None else "the N = 1 point of this workload is `secondary.supremacy` of the "
                                            "`--gpus 1` line (its primary workload is QFT-33 complex128)",
                       "l2_policy": "shards are far larger than the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the launch stream, max over ranks"},
            "roofline": primary["roofline"],
            "parity": primary["parity"],
            "e2e": e2e,
            "gpu_launches": primary["gpu_launches"],
            "clocks": primary["clocks"],
        }
        if secondary:
            line["secondary"] = secondary
        if "measurement" in primary:
            line["measurement"] = primary["measurement"]
    dist.barrier()
    dist.destroy_process_group()
    return line          # rank 0: the JSON line (bench.py writes it to the real stdout); other ranks: None
