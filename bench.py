#!/usr/bin/env python
"""bench.py -- circuit throughput of the B200 state-vector gate path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload variational|qft|supremacy|qv]
                    [--nqubits n] [--dtype complex128|complex64] [--fuse k] [--impl reference]

A "step" is one full execution of the workload circuit: |0..0> preparation followed by every
gate (fused into blocks of at most --fuse qubits).  The default workload is BASELINE.json's
configs[1]: the variational RY+CZ circuit on 30 qubits in complex128 with gate fusion, on one
B200.  `value` is gates/s with the circuit program resident (state and kernels on the device,
gate matrices already built); `e2e` is the same metric through the public backend API
(`execute_circuit` from host gate objects: host matrices go down with every launch, a marginal
probability vector comes back every step).  `roofline` is measured live with CUDA events around
every launch of the timed steps; `cpu_baseline` times the CPU oracle (a C/OpenMP port of the
reference's numba kernels) on this box's host cores on a bounded sample of the same gate list.

`--impl reference` runs that CPU port alone (the reference itself is Python+numba+qibo and
cannot be installed offline) and prints the same JSON line.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEFAULTS = {
    "variational": dict(nqubits=30, dtype="complex128", fuse=4),
    "qft": dict(nqubits=33, dtype="complex128", fuse=1),
    "supremacy": dict(nqubits=32, dtype="complex64", fuse=4),
    "qv": dict(nqubits=32, dtype="complex64", fuse=2),
}


def build_circuit(workload, nqubits):
    from qibojit_b200 import circuits

    if workload == "variational":
        return circuits.variational(nqubits)
    if workload == "qft":
        return circuits.qft(nqubits)
    if workload == "supremacy":
        return circuits.supremacy(nqubits)
    if workload == "qv":
        return circuits.quantum_volume(nqubits)
    raise ValueError(workload)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                     "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU arm
def oracle_program(circuit):
    """Lower gate objects to calls of the CPU oracle (same dispatch as the reference backend)."""
    from oracle import oracle as O
    from qibojit_b200.backends.b200 import GATE_OPS
    from qibojit_b200.matrices import CustomMatrices
    from tests import refdispatch as R

    n = circuit.nqubits
    prog = []
    for g in circuit.queue:
        name = g.__class__.__name__
        t = g.target_qubits
        q = R.qubits_tensor(n, t, g.control_qubits)

        def mat(dtype, g=g, name=name):
            from qibojit_b200 import fusion
            mats = CustomMatrices(dtype)
            if name == "FusedGate":
                return fusion.fused_matrix(g, mats)
            return g.target_matrix(mats)

        if len(t) == 1:
            op = GATE_OPS.get(name, "apply_gate")
            prog.append(lambda st, dt, t=t, op=op, mat=mat, q=q, g=g: R.one_qubit_base(
                O, st, n, t[0], op, mat(dt), q if g.control_qubits else None))
        elif len(t) == 2:
            op = GATE_OPS.get(name, "apply_two_qubit_gate")
            prog.append(lambda st, dt, t=t, op=op, mat=mat, q=q, g=g: R.two_qubit_base(
                O, st, n, t[0], t[1], op, mat(dt), q if g.control_qubits else None))
        else:
            prog.append(lambda st, dt, t=t, mat=mat, q=q: R.multi_qubit_base(O, st, n, list(t), mat(dt), q))
    return prog


def cpu_sample(workload, nqubits, dtype, budget_s, fuse):
    """Time the oracle on a bounded sample: the first gates of the same (unfused, as the
    reference executes it by default) circuit at the largest n <= nqubits that fits host RAM,
    until `budget_s` seconds are spent.  Returns (gates/s, description, cores)."""
    import psutil

    from oracle import oracle as O

    amp = 16 if dtype == "complex128" else 8
    avail = psutil.virtual_memory().available
    n = nqubits
    while (amp << n) > 0.6 * avail and n > 20:
        n -= 1
    circuit = build_circuit(workload, n)
    if fuse > 1:  # qibo's default fusion width for the numba backend is two qubits
        circuit = circuit.fuse(max_qubits=2)
    weights = [len(getattr(g, "gates", [g])) for g in circuit.queue]
    total_gates = sum(weights)
    prog = oracle_program(circuit)
    cores = O.max_threads()
    st = np.empty(1 << n, dtype=dtype)
    O.initial_state_vector(st)
    prog[0](st, dtype)  # warm up (page faults, thread pool)
    O.initial_state_vector(st)
    t0 = time.perf_counter()
    done = 0
    for call, w in zip(prog, weights):
        call(st, dtype)
        done += w
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    gps = done / dt
    how = "fused to 2-qubit blocks as qibo does by default" if fuse > 1 else "unfused"
    desc = (f"first {done} of {total_gates} gates of {workload}-{n} {dtype} ({how}) in {dt:.1f}s"
            f" on {cores} threads")
    if n != nqubits:
        scale = 2.0 ** (n - nqubits)
        desc += f"; n reduced from {nqubits} to fit host RAM, value scaled by 2^{n - nqubits}"
        gps *= scale
    return gps, desc, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = DEFAULTS[args.workload]
    nqubits = args.nqubits or cfg["nqubits"]
    dtype = args.dtype or cfg["dtype"]
    t0 = time.perf_counter()
    per_step = max(5.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    desc, cores = "", 1
    for i in range(args.warmup + args.steps):
        gps, desc, cores = cpu_sample(args.workload, nqubits, dtype, per_step, cfg["fuse"])
        if i >= args.warmup:
            vals.append(gps)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "gates_per_second", "value": value, "unit": "gates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (time.perf_counter() - t0) / max(1, args.steps + args.warmup),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if dtype == "complex128" else "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}-{nqubits}-{dtype}", "fusion_max_qubits": min(2, cfg["fuse"]),
                   "note": "CPU port (oracle/qj_oracle.c, C+OpenMP) of the reference numba kernels"},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def measurement_leg(backend, state, nqubits, dtype, nshots=10 ** 6):
    """BASELINE.json configs[4]: full-register probabilities, 10^6-shot sampling (the reference's
    Metropolis sampler semantics, ops.py:86-108) and one collapse on three qubits, timed once each
    with CUDA events on the launch stream.  Checked: the shot counts add up, the collapsed state is
    normalised."""
    import torch

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1)

    amp = 16 if dtype == "complex128" else 8
    np.random.seed(123)
    probs, ms_probs = timed(lambda: backend.calculate_probabilities(state, list(range(nqubits)), nqubits))
    freqs, ms_sample = timed(lambda: backend.sample_frequencies(probs, nshots))
    assert sum(freqs.values()) == nshots, sum(freqs.values())
    del probs
    torch.cuda.empty_cache()
    shot = max(freqs, key=freqs.get)
    qubits = [1, nqubits // 2, nqubits - 2]
    outcome = sum(((shot >> (nqubits - 1 - q)) & 1) << (len(qubits) - 1 - i) for i, q in enumerate(qubits))
    _, ms_collapse = timed(lambda: backend.collapse_state(state, qubits, outcome, nqubits))
    norm = backend.calculate_norm(state)
    assert abs(norm - 1.0) < (1e-9 if dtype == "complex128" else 1e-4), norm
    nbytes = amp << nqubits
    return {"nshots": nshots, "distinct_outcomes": len(freqs),
            "probabilities_ms": ms_probs, "probabilities_gbs": nbytes / (ms_probs * 1e-3) / 1e9,
            "sample_frequencies_ms": ms_sample, "shots_per_second": nshots / (ms_sample * 1e-3),
            "collapse_qubits": qubits, "collapse_ms": ms_collapse,
            # zero 7/8 of the state, read 1/8 for the norm, rescale 1/8
            "collapse_gbs": nbytes * (7 / 8 + 1 / 8 + 2 / 8) / (ms_collapse * 1e-3) / 1e9,
            "norm_after_collapse": norm}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries the one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from qibojit_b200 import _capi
    from qibojit_b200.backends.b200 import B200Backend

    cfg = DEFAULTS[args.workload]
    nqubits = args.nqubits or cfg["nqubits"]
    dtype = args.dtype or cfg["dtype"]
    fuse = args.fuse or cfg["fuse"]
    backend = B200Backend()
    backend.set_dtype(dtype)

    if world > 1:
        from bench_distributed import run_distributed

        return run_distributed(args, backend, nqubits, dtype, fuse, world, rank)

    circuit = build_circuit(args.workload, nqubits)
    ngates = circuit.ngates
    amp = 16 if dtype == "complex128" else 8
    nbytes_state = amp << nqubits
    lib, h = backend._lib, backend._handle()

    # the circuit program: multi-gate passes (k_pass) + the gates the planner leaves to the
    # per-gate kernels.  Compiled once, resident on the device (the `value` leg).
    t0 = time.perf_counter()
    prog = backend.compile_circuit(circuit, zero_state=True)   # every step starts from |0...0>, as execute_circuit does
    plan_ms = 1e3 * (time.perf_counter() - t0)
    pstats = prog.stats()
    state = backend.zero_state(nqubits)
    tag = backend._tag(state)

    def step(events=None):
        nonlocal state
        _capi.check(lib.qj_initial_state(h, state.data_ptr(), tag, nqubits))
        if events is None:
            state = prog.run(state)
            return

        def timer(kind, frac, fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            events.append((kind, 2.0 * nbytes_state * frac, e0, e1))

        state = prog.run_timed(state, timer)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = backend.launch_count()
    events = []
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        t_start.record()
        for _ in range(args.steps):
            step(events)
        t_end.record()
        torch.cuda.synchronize()
    total_ms = t_start.elapsed_time(t_end)
    launches = backend.launch_count() - launches0
    ms_per_step = total_ms / args.steps
    value = ngates / (ms_per_step * 1e-3)

    # roofline of the dominant kernel class, from the per-launch events of the timed steps
    per_kind = {}
    for kind, alg, e0, e1 in events:
        d = per_kind.setdefault(kind, {"ms": 0.0, "bytes": 0.0, "n": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["bytes"] += alg
        d["n"] += 1
    dom = max(per_kind, key=lambda k: per_kind[k]["ms"])
    peak, peak_src = measured_peak_gbs()
    achieved = per_kind[dom]["bytes"] / (per_kind[dom]["ms"] * 1e-3) / 1e9
    breakdown = {k: {"launches_per_step": v["n"] // args.steps, "avg_ms": v["ms"] / v["n"],
                     "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9} for k, v in per_kind.items()}
    kernel_names = {"pass": "k_pass (multi-gate tile pass, 2*N*A bytes per launch)"}
    # arithmetic side: a pass that absorbs many dense gates is bound by the FP64 / FP32 pipe, not by HBM
    fma = prog.fma_per_amplitude()
    pass_ms = per_kind.get("pass", {"ms": 0.0})["ms"] / args.steps
    fp_peak = 37.2 if dtype == "complex128" else 74.5   # nominal CUDA-core peaks: 148 SMs x 64 (128) lanes x 2 x 1.965 GHz
    fp = {"fma_per_amplitude_per_step": fma, "tflops": 2.0 * fma * 2.0 ** nqubits / (pass_ms * 1e-3) / 1e12 if pass_ms else None,
          "peak_tflops_nominal": fp_peak, "pipe": "fp64" if dtype == "complex128" else "fp32",
          "hbm_bound_ms": pstats["launches"] * 2.0 * nbytes_state / (peak * 1e9) * 1e3,
          "fp_bound_ms": 2.0 * fma * 2.0 ** nqubits / (fp_peak * 1e12) * 1e3}
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "r1o_ncu_pass_var30_full_summary.csv")
    if args.workload == "variational" and nqubits == 30 and dtype == "complex128" and os.path.exists(prof):
        import csv

        vals = {r[0]: r[2] for r in csv.reader(open(prof)) if len(r) == 3}
        traffic = (float(vals["dram__bytes_read.sum"]) + float(vals["dram__bytes_write.sum"])) * 1e9
        traffic_src = "profiles/r1o_ncu_pass_var30_full_summary.csv (ncu --set full, first k_pass launch of this workload)"

    measurement = None
    if args.workload == "qv" or args.measure:
        measurement = measurement_leg(backend, state, nqubits, dtype)

    # end to end through the public API, every step from HOST gate objects: plan + encode the
    # passes, upload the program image and phase tables, run, read a marginal back
    del state  # QFT-33 fills 137 GB of the 180 GB: the e2e leg allocates its own state
    torch.cuda.empty_cache()
    e2e_times = []
    d2h = 0
    h2d = 0
    for i in range(1 + min(args.steps, 3)):
        circuit.__dict__.pop("_qj_programs", None)   # no cached program: compile inside the timed region
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = backend.execute_circuit(circuit)
        probs = backend.calculate_probabilities(out, [0, 1, 2, 3], nqubits)
        host = probs.cpu().numpy()
        torch.cuda.synchronize()
        if i:
            e2e_times.append(time.perf_counter() - t0)
        d2h = host.nbytes
        h2d = sum(entry[1].upload_bytes for entry in circuit.__dict__.get("_qj_programs", {}).values())
        del out, probs      # the block goes back to torch's caching allocator and is reused by the next step
    e2e_value = ngates / float(np.mean(e2e_times))
    assert abs(host.sum() - 1.0) < (1e-6 if dtype == 'complex128' else 1e-3), host.sum()

    cpu_gps, cpu_desc, cores = cpu_sample(args.workload, nqubits, dtype, args.cpu_seconds, fuse)

    line = {
        "metric": "gates_per_second", "value": value, "unit": "gates/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if dtype == "complex128" else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-{nqubits}-{dtype}", "circuit_gates": ngates,
                   "execution": "multi-gate tile passes (planner.Program compiled for the |0...0> input: SWAP gates are relabellings)",
                   "passes": pstats["passes"], "launches_per_step": pstats["launches"] + pstats["raw_gates"] + 1,
                   "rounds": pstats["rounds"], "micro_ops": pstats["micro_ops"], "raw_gates": pstats["raw_gates"],
                   "plan_compile_ms": plan_ms, "state_bytes": nbytes_state,
                   "l2_policy": "state (>= 16 GiB) is far larger than the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the launch stream"},
        "roofline": {"bound": "hbm", "kernel": kernel_names.get(dom, dom), "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": 2.0 * nbytes_state, "peak_source": peak_src,
                     "per_kernel": breakdown, "arithmetic": fp},
        "cpu_baseline": {"value": cpu_gps, "unit": "gates/s", "cores": cores, "kind": "port", "sample": cpu_desc},
        "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h),
                "includes": "planning, program encode + upload, state preparation, all passes, marginal read-back"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if measurement is not None:
        line["measurement"] = measurement
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="variational", choices=sorted(DEFAULTS))
    ap.add_argument("--nqubits", type=int, default=0)
    ap.add_argument("--dtype", default="")
    ap.add_argument("--fuse", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--measure", action="store_true",
                    help="append the measurement leg (probabilities, 10^6 shots, collapse); default for --workload qv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
