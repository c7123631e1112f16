#!/usr/bin/env python
"""bench.py -- circuit throughput of the B200 state-vector gate path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload qft|variational|supremacy|qv]
                    [--nqubits n] [--dtype complex128|complex64] [--impl reference] [--no-secondary]

A "step" is one full execution of the workload circuit: |0..0> preparation followed by every gate.

Defaults (BASELINE.json north-star configs):
  * N = 1: QFT on 33 qubits, complex128 (137 GB of state, configs[2]) -- the primary line; the
    variational-30 circuit (configs[1]) and the supremacy-34 complex64 circuit (the N = 1 point of
    the multi-GPU series) ride along as `secondary` records of the same JSON line;
  * N > 1: the supremacy-style circuit on 34 qubits, complex64, sharded over the ranks (strong
    scaling: 34 qubits fit one GPU, so the N = 1 point exists -- it is `secondary.supremacy` of the
    N = 1 line); QFT-33 sharded and, on 8 ranks, supremacy-36 (configs[3]) are `secondary` records.

`value` is gates/s with the circuit program resident (state and kernels on the device);
`e2e` is the same metric through the public backend API (`execute_circuit` from host gate objects:
planning, program encoding and upload inside the timed region, a marginal probability vector read
back every step).  `roofline` is measured live with CUDA events around every launch of the timed
steps.  Parity is checked inside the run at the benchmark size: QFT|0..0> against its closed form
(max |amp - 2^(-n/2)| on the device), the other circuits' 4-qubit marginals against committed
fixtures made with the reference's numba kernels (tests/golden/make_marginals.py); a mismatch
above 1e-12 (complex128) / 1e-5 (complex64) fails the run.

`cpu_baseline` / `--impl reference` time the REFERENCE's own numba kernels (oracle/_ref, see
oracle/make_ref.py) on this box's host cores -- thread count = CPU affinity, as the reference backend
sets it (backends/cpu.py:86-89) whatever OMP_NUM_THREADS says -- on a bounded sample of the same gate
list; when oracle/_ref is absent the C/OpenMP port (oracle/qj_oracle.c) stands in (`kind: "port"`).
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
    "qft": dict(nqubits=33, dtype="complex128", cpu_fuse=False),
    "variational": dict(nqubits=30, dtype="complex128", cpu_fuse=True),
    "supremacy": dict(nqubits=34, dtype="complex64", cpu_fuse=True),
    "qv": dict(nqubits=32, dtype="complex64", cpu_fuse=True),
}
TOL = {"complex128": 1e-12, "complex64": 1e-5}
MARGINAL_QUBITS = [0, 1, 2, 3]
FP_PEAK_TFLOPS = {"complex128": 37.2, "complex64": 74.5}   # nominal CUDA-core peaks: 148 SMs x 64 (128) lanes x 2 x 1.965 GHz


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


def marginal_fixture(workload, nqubits, dtype):
    """Committed 4-qubit marginal of `workload` (tests/golden/marginals.json), or None."""
    path = os.path.join(ROOT, "tests", "golden", "marginals.json")
    try:
        with open(path) as f:
            entry = json.load(f).get(f"{workload}-{nqubits}-{dtype}")
    except Exception:
        return None
    if entry is None:
        return None
    return np.asarray(entry["marginal"], dtype=np.float64), entry["source"]


def ncu_traffic(workload, nqubits, dtype):
    """dram__bytes_read + dram__bytes_write per k_pass launch from the committed ncu capture of this
    workload (profiles/round2_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep)."""
    path = os.path.join(ROOT, "profiles", "round2_traffic.json")
    try:
        with open(path) as f:
            entry = json.load(f).get(f"{workload}-{nqubits}-{dtype}")
    except Exception:
        return None, None
    if entry is None:
        return None, None
    return float(entry["bytes_per_launch"]), entry["source"]


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
def host_threads():
    """The thread count the reference backend uses: the CPU affinity of the process
    (backends/cpu.py:86-89), whatever OMP_NUM_THREADS (torchrun sets it to 1) says."""
    try:
        import psutil

        return len(psutil.Process().cpu_affinity())
    except Exception:
        return os.cpu_count() or 1


def cpu_kernels():
    """-> (gate kernels, ops kernels, threads, kind): the reference's numba modules from oracle/_ref
    ('reference') or, when that build output is missing, the C/OpenMP port ('port')."""
    n = host_threads()
    try:
        from oracle import numba_ref

        G, O, used = numba_ref.load(n)
        return G, O, used, "reference"
    except Exception as exc:                      # oracle/_ref absent or numba unusable
        from oracle import oracle as P

        P.set_threads(n)
        sys.stderr.write(f"bench: reference numba kernels unavailable ({exc}); timing the C port\n")
        return P, P, P.max_threads(), "port"


def reference_program(circuit, K):
    """Lower gate objects to calls of the kernel module `K` with the reference backend's dispatch
    (cpu.py:433-450, 519-635 restated in tests/refdispatch.py)."""
    from qibojit_b200 import fusion
    from qibojit_b200.backends.b200 import GATE_OPS
    from qibojit_b200.matrices import CustomMatrices
    from tests import refdispatch as R

    n = circuit.nqubits
    prog = []
    cache = {}

    def mats(dtype):
        if dtype not in cache:
            cache[dtype] = CustomMatrices(dtype)
        return cache[dtype]

    for g in circuit.queue:
        name = g.__class__.__name__
        t = g.target_qubits
        q = R.qubits_tensor(n, t, g.control_qubits)

        def mat(dtype, g=g, name=name):
            if name == "FusedGate":
                return np.ascontiguousarray(fusion.fused_matrix(g, mats(dtype)))
            return g.target_matrix(mats(dtype))

        if len(t) == 1:
            op = GATE_OPS.get(name, "apply_gate")
            prog.append(lambda st, dt, t=t, op=op, mat=mat, q=q, g=g: R.one_qubit_base(
                K, st, n, t[0], op, mat(dt), q if g.control_qubits else None))
        elif len(t) == 2:
            op = GATE_OPS.get(name, "apply_two_qubit_gate")
            prog.append(lambda st, dt, t=t, op=op, mat=mat, q=q, g=g: R.two_qubit_base(
                K, st, n, t[0], t[1], op, mat(dt), q if g.control_qubits else None))
        else:
            prog.append(lambda st, dt, t=t, mat=mat, q=q: R.multi_qubit_base(K, st, n, list(t), mat(dt), q))
    return prog


def cpu_sample(workload, nqubits, dtype, budget_s, fuse, full_if_within=None):
    """Time the CPU kernels on a bounded sample: the first gates of the same circuit (fused to
    two-qubit blocks when `fuse`, as qibo does by default for the numba backend) at the largest
    n <= nqubits that fits host RAM, until `budget_s` seconds are spent.  Every kernel signature is
    compiled (numba JIT) before the clock starts, as the reference's benchmark does with its dry run
    (benchmarks/main.py:77-94).  Returns a dict (value in gates/s scaled to `nqubits`)."""
    import psutil

    G, O, cores, kind = cpu_kernels()
    amp = 16 if dtype == "complex128" else 8
    avail = psutil.virtual_memory().available
    n = nqubits
    while (amp << n) > 0.6 * avail and n > 20:
        n -= 1
    circuit = build_circuit(workload, n)
    if fuse:
        circuit = circuit.fuse(max_qubits=2)
    weights = [len(getattr(g, "gates", [g])) for g in circuit.queue]
    total_gates = sum(weights)
    # JIT warm-up of every kernel signature on a small register of the same circuit family
    small = build_circuit(workload, 12)
    if fuse:
        small = small.fuse(max_qubits=2)
    st = np.empty(1 << 12, dtype=dtype)
    O.initial_state_vector(st)
    for call in reference_program(small, G):
        call(st, dtype)
    prog = reference_program(circuit, G)
    st = np.empty(1 << n, dtype=dtype)
    O.initial_state_vector(st)
    prog[0](st, dtype)  # page faults, thread pool
    O.initial_state_vector(st)
    t0 = time.perf_counter()
    done = 0
    ncalls = 0
    for call, w in zip(prog, weights):
        call(st, dtype)
        done += w
        ncalls += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    gps = done / dt
    how = "fused to 2-qubit blocks as qibo does by default" if fuse else "unfused"
    what = "reference numba kernels (oracle/_ref)" if kind == "reference" else "C/OpenMP port of the numba kernels"
    desc = (f"first {done} of {total_gates} gates of {workload}-{n} {dtype} ({how}) in {dt:.1f}s"
            f" on {cores} threads, {what}")
    if n != nqubits:
        scale = 2.0 ** (n - nqubits)
        desc += f"; n reduced from {nqubits} to fit host RAM, value scaled by 2^{n - nqubits}"
        gps *= scale
    out = {"value": gps, "unit": "gates/s", "cores": cores, "kind": kind, "sample": desc}
    if ncalls == len(prog) and n == nqubits:
        out["_final_state"] = st        # the whole circuit ran: its marginal is a parity reference
    return out


def qft20_cpu():
    """BASELINE.json configs[0]: QFT on 20 qubits, complex128, on the CPU kernels (second run timed)."""
    G, O, cores, kind = cpu_kernels()
    n = 20
    circuit = build_circuit("qft", n)
    prog = reference_program(circuit, G)
    best = None
    st = None
    for _ in range(3):
        st = np.empty(1 << n, dtype=np.complex128)
        O.initial_state_vector(st)
        t0 = time.perf_counter()
        for call in prog:
            call(st, "complex128")
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    err = float(np.abs(st - 2.0 ** (-n / 2)).max())
    return {"workload": "qft-20-complex128", "backend": f"CPU, {kind}", "cores": cores, "gates": circuit.ngates,
            "ms": 1e3 * best, "gates_per_s": circuit.ngates / best, "max_abs_err_vs_closed_form": err}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload or ("qft" if args.gpus == 1 else "supremacy")
    cfg = DEFAULTS[workload]
    nqubits = args.nqubits or cfg["nqubits"]
    dtype = args.dtype or cfg["dtype"]
    circuit_gates = build_circuit(workload, nqubits).ngates
    per_step = max(5.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    sample = None
    for i in range(args.warmup + args.steps):
        sample = cpu_sample(workload, nqubits, dtype, per_step, cfg["cpu_fuse"])
        sample.pop("_final_state", None)
        if i >= args.warmup:
            vals.append(sample["value"])
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "gates_per_second", "value": value, "unit": "gates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # the time one full circuit takes at this rate (each step timed a bounded sample of it)
        "ms_per_step": 1e3 * circuit_gates / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if dtype == "complex128" else "f32", "data": "synthetic",
        "config": {"workload": f"{workload}-{nqubits}-{dtype}", "circuit_gates": circuit_gates,
                   "fusion_max_qubits": 2 if cfg["cpu_fuse"] else 1,
                   "note": "the reference's CPU implementation of the path on this box's host cores; "
                           "ms_per_step = circuit_gates / value (a step times a bounded sample)"},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": sample["cores"], "kind": sample["kind"],
                         "sample": sample["sample"]},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        line["configs0"] = qft20_cpu()
    except Exception as exc:
        line["configs0"] = {"error": str(exc)}
    emit(json.dumps(line))


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
            # |psi|^2 reads N*A bytes and writes N*A/2 (one real per amplitude)
            "probabilities_ms": ms_probs, "probabilities_gbs": 1.5 * nbytes / (ms_probs * 1e-3) / 1e9,
            "sample_frequencies_ms": ms_sample, "shots_per_second": nshots / (ms_sample * 1e-3),
            "collapse_qubits": qubits, "collapse_ms": ms_collapse,
            # zero 7/8 of the state, read 1/8 for the norm, rescale 1/8
            "collapse_gbs": nbytes * (7 / 8 + 1 / 8 + 2 / 8) / (ms_collapse * 1e-3) / 1e9,
            "norm_after_collapse": norm}


def check_parity(backend, state, workload, nqubits, dtype, cpu_state=None):
    """Parity inside the run, at the benchmark size.  Raises on a mismatch above TOL[dtype]."""
    tol = TOL[dtype]
    out = {"tolerance": tol}
    if workload == "qft":
        err = backend.max_deviation(state, 2.0 ** (-nqubits / 2))
        out.update(check="max |amp - 2^(-n/2)| over all 2^n amplitudes (closed form of QFT|0..0>), on the device",
                   max_abs_err=err)
        if not err <= tol:
            raise AssertionError(f"parity: qft-{nqubits} {dtype} deviates from 2^(-n/2) by {err:.3e} > {tol}")
        return out
    marg = backend.calculate_probabilities(state, MARGINAL_QUBITS, nqubits).double().cpu().numpy()
    out["marginal_sum"] = float(marg.sum())
    ref, src = None, None
    if cpu_state is not None:
        p = (np.abs(cpu_state.astype(np.complex128)) ** 2).reshape((2,) * 4 + (-1,)).sum(axis=-1).reshape(-1)
        ref, src = p, "the CPU leg of this run (reference kernels, whole circuit)"
    else:
        fx = marginal_fixture(workload, nqubits, dtype)
        if fx is not None:
            ref, src = fx
    if ref is None:
        out.update(check="4-qubit marginal sums to 1 (no reference marginal at this size)", pinned=False)
        if abs(out["marginal_sum"] - 1.0) > (1e-9 if dtype == "complex128" else 1e-4):
            raise AssertionError(f"parity: marginal of {workload}-{nqubits} sums to {out['marginal_sum']}")
        return out
    err = float(np.abs(marg - ref).max())
    out.update(check=f"4-qubit marginal (qubits {MARGINAL_QUBITS}) vs {src}", max_abs_err=err, pinned=True)
    if not err <= tol:
        raise AssertionError(f"parity: marginal of {workload}-{nqubits} {dtype} differs by {err:.3e} > {tol}")
    return out


def time_program(backend, workload, nqubits, dtype, steps, warmup, zero_state=True, sampler=None):
    """Compile the circuit into multi-gate passes and time `steps` executions (|0..0> preparation +
    every launch), each launch bracketed by CUDA events on the launch stream.  Returns the record
    and the final state (for the parity check)."""
    import torch

    from qibojit_b200 import _capi

    circuit = build_circuit(workload, nqubits)
    amp = 16 if dtype == "complex128" else 8
    nbytes_state = amp << nqubits
    backend.set_dtype(dtype)
    t0 = time.perf_counter()
    prog = backend.compile_circuit(circuit, zero_state=zero_state)
    plan_ms = 1e3 * (time.perf_counter() - t0)
    pstats = prog.stats()
    state = backend.zero_state(nqubits)
    events = []

    def step(timed):
        # every step starts from |0...0>: the first pass prepares it (it writes every amplitude
        # without reading any), as execute_circuit does
        nonlocal state
        if not timed:
            state = prog.run(state, from_zero=True)
            return

        def timer(kind, frac, fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            events.append((kind, 2.0 * nbytes_state * frac, e0, e1))

        state = prog.run_timed(state, timer, from_zero=True)

    for _ in range(warmup):
        step(False)
    torch.cuda.synchronize()
    launches0 = backend.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = sampler if sampler is not None else ClockSampler(torch.cuda.current_device())
    with ctx as clocks:
        torch.cuda.synchronize()
        t_start.record()
        for _ in range(steps):
            step(True)
        t_end.record()
        torch.cuda.synchronize()
    total_ms = t_start.elapsed_time(t_end)
    launches = backend.launch_count() - launches0
    ms_per_step = total_ms / steps

    peak, peak_src = measured_peak_gbs()
    per_kind = {}
    for kind, alg, e0, e1 in events:
        d = per_kind.setdefault(kind, {"ms": 0.0, "bytes": 0.0, "n": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["bytes"] += alg
        d["n"] += 1
    dom = max(per_kind, key=lambda k: per_kind[k]["ms"])
    achieved = per_kind[dom]["bytes"] / (per_kind[dom]["ms"] * 1e-3) / 1e9
    breakdown = {k: {"launches_per_step": v["n"] // steps, "avg_ms": v["ms"] / v["n"],
                     "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9} for k, v in per_kind.items()}
    # per pass: time, algorithmic GB/s and the fraction of the bound that applies to it -- the slower
    # of HBM (2*N*A bytes at the measured copy peak) and the FP pipe (its multiply-adds at the
    # nominal FP64 / FP32 CUDA-core peak): a pass that absorbs many dense gates is FP-bound
    pass_events = [(e0.elapsed_time(e1), alg) for kind, alg, e0, e1 in events if kind in ("pass", "pass0")]
    npass = len(pass_events) // steps if steps else 0
    fma_pass = prog.fma_per_pass()
    fp_peak = FP_PEAK_TFLOPS[dtype]
    per_pass, bound_ms, spent_ms = [], 0.0, 0.0
    for i in range(npass):
        ms = float(np.mean([pass_events[s * npass + i][0] for s in range(steps)]))
        alg = pass_events[i][1]
        hbm_ms = alg / (peak * 1e9) * 1e3
        rec = {"ms": ms, "gbs": alg / (ms * 1e-3) / 1e9, "frac_hbm": hbm_ms / ms}
        if len(fma_pass) == npass:
            fp_ms = 2.0 * fma_pass[i] * 2.0 ** nqubits / (fp_peak * 1e12) * 1e3
            rec.update(fma_per_amplitude=fma_pass[i], fp_bound_ms=fp_ms, hbm_bound_ms=hbm_ms,
                       frac_of_bound=max(hbm_ms, fp_ms) / ms)
            bound_ms += max(hbm_ms, fp_ms)
        else:
            bound_ms += hbm_ms
        spent_ms += ms
        per_pass.append(rec)
    fma = prog.fma_per_amplitude()
    pass_ms = (per_kind.get("pass", {"ms": 0.0})["ms"] + per_kind.get("pass0", {"ms": 0.0})["ms"]) / steps
    all_bytes = sum(v["bytes"] for k, v in per_kind.items() if k in ("pass", "pass0"))
    all_ms = sum(v["ms"] for k, v in per_kind.items() if k in ("pass", "pass0"))
    traffic, traffic_src = ncu_traffic(workload, nqubits, dtype)
    record = {
        "workload": f"{workload}-{nqubits}-{dtype}", "circuit_gates": circuit.ngates,
        "ms_per_step": ms_per_step, "value": circuit.ngates / (ms_per_step * 1e-3),
        "compiled_for": "the |0...0> input (uncontrolled SWAP gates become relabellings)" if zero_state
                        else "any input state (SWAP gates move data)",
        "passes": pstats["passes"], "launches_per_step": pstats["launches"] + pstats["raw_gates"],
        "rounds": pstats["rounds"], "micro_ops": pstats["micro_ops"], "raw_gates": pstats["raw_gates"],
        "plan_compile_ms": plan_ms, "state_bytes": nbytes_state, "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "kernel": "k_pass (multi-gate tile pass, 2*N*A bytes per launch)" if dom == "pass" else dom,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "launches_in_frac": "the k_pass launches that read AND write the state (per_kernel.pass); the first launch of a "
                                "step starts from |0...0>, only writes (N*A algorithmic bytes, state preparation fused in) "
                                "and is per_kernel.pass0 / per_pass[0]",
            "all_pass_launches": {"achieved": all_bytes / (all_ms * 1e-3) / 1e9 if all_ms else None,
                                  "frac": all_bytes / (all_ms * 1e-3) / 1e9 / peak if all_ms else None},
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": 2.0 * nbytes_state, "peak_source": peak_src,
            "per_kernel": breakdown, "per_pass": per_pass,
            "frac_of_max_hbm_fp_bound": (bound_ms / spent_ms) if spent_ms else None,
            "arithmetic": {"fma_per_amplitude_per_step": fma,
                           "tflops": 2.0 * fma * 2.0 ** nqubits / (pass_ms * 1e-3) / 1e12 if pass_ms else None,
                           "peak_tflops_nominal": fp_peak, "pipe": "fp64" if dtype == "complex128" else "fp32"},
        },
        "clocks": clocks.summary(),
    }
    return record, state, circuit


def time_e2e(backend, circuit, nqubits, dtype, reps):
    """End to end through the public API, every step from HOST gate objects: plan + encode the
    passes, upload the program images and phase tables, prepare the state, run, read a marginal back."""
    import torch

    times, d2h, h2d, host = [], 0, 0, None
    for i in range(1 + reps):
        circuit.__dict__.pop("_qj_programs", None)   # no cached program: compile inside the timed region
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = backend.execute_circuit(circuit)
        probs = backend.calculate_probabilities(out, MARGINAL_QUBITS, nqubits)
        host = probs.cpu().numpy()
        torch.cuda.synchronize()
        if i:
            times.append(time.perf_counter() - t0)
        d2h = host.nbytes
        h2d = sum(entry[1].upload_bytes for entry in circuit.__dict__.get("_qj_programs", {}).values())
        del out, probs      # the block goes back to torch's caching allocator and is reused by the next step
    assert abs(host.sum() - 1.0) < (1e-6 if dtype == "complex128" else 1e-3), host.sum()
    return {"value": circuit.ngates / float(np.mean(times)), "unit": "gates/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(np.mean(times)),
            "includes": "planning, program encode + upload, state preparation, all passes, marginal read-back "
                        "(the final state stays on the device, as CupyBackend leaves it; it is not copied to the host)"}


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

    from qibojit_b200.backends.b200 import B200Backend

    backend = B200Backend()
    if world > 1:
        from bench_distributed import run_distributed

        line = run_distributed(args, backend, world, rank)
        if line is not None:
            emit(json.dumps(line))
        return

    workload = args.workload or "qft"
    cfg = DEFAULTS[workload]
    nqubits = args.nqubits or cfg["nqubits"]
    dtype = args.dtype or cfg["dtype"]

    # ---- primary workload: resident program
    primary, state, circuit = time_program(backend, workload, nqubits, dtype, args.steps, args.warmup, zero_state=True)
    parity = check_parity(backend, state, workload, nqubits, dtype)
    measurement = None
    if workload == "qv" or args.measure:
        measurement = measurement_leg(backend, state, nqubits, dtype)
    del state
    torch.cuda.empty_cache()

    # the same circuit compiled for an arbitrary input state (the SWAP gates move data); run from
    # |0..0> so that the closed form still checks it
    general = None
    if workload == "qft":
        rec, st, _ = time_program(backend, workload, nqubits, dtype, min(args.steps, 3), 1, zero_state=False)
        gp = check_parity(backend, st, workload, nqubits, dtype)
        general = {k: rec[k] for k in ("ms_per_step", "value", "passes", "compiled_for")}
        general["roofline_frac"] = rec["roofline"]["frac"]
        general["parity_max_abs_err"] = gp.get("max_abs_err")
        del st
        torch.cuda.empty_cache()

    # ---- end to end through the public API
    e2e = time_e2e(backend, circuit, nqubits, dtype, min(args.steps, 3))
    torch.cuda.empty_cache()

    # ---- secondary workloads (other BASELINE configs), same bar: timed, roofline, parity
    secondary = {}
    if not args.no_secondary and not args.workload:
        for name in ("variational", "supremacy"):
            c = DEFAULTS[name]
            try:
                rec, st, _ = time_program(backend, name, c["nqubits"], c["dtype"], min(args.steps, 3), 3)
                rec["parity"] = check_parity(backend, st, name, c["nqubits"], c["dtype"])
                del st
            except AssertionError:
                raise
            except Exception as exc:          # (e.g. not enough memory left on a shared device)
                rec = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
            secondary[name] = rec
    backend.set_dtype(dtype)

    # ---- CPU baseline: the reference's own numba kernels on this box's host cores
    cpu = cpu_sample(workload, nqubits, dtype, args.cpu_seconds, cfg["cpu_fuse"])
    cpu.pop("_final_state", None)
    try:
        configs0 = qft20_cpu()
    except Exception as exc:
        configs0 = {"error": str(exc)}

    line = {
        "metric": "gates_per_second", "value": primary["value"], "unit": "gates/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": primary["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if dtype == "complex128" else "f32", "data": "synthetic",
        "config": {"workload": primary["workload"], "circuit_gates": primary["circuit_gates"],
                   "execution": "multi-gate tile passes (planner.Program), " + primary["compiled_for"],
                   "passes": primary["passes"], "launches_per_step": primary["launches_per_step"],
                   "rounds": primary["rounds"], "micro_ops": primary["micro_ops"], "raw_gates": primary["raw_gates"],
                   "plan_compile_ms": primary["plan_compile_ms"], "state_bytes": primary["state_bytes"],
                   "l2_policy": "state (>= 16 GiB) is far larger than the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the launch stream"},
        "roofline": primary["roofline"],
        "parity": parity,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": primary["gpu_launches"],
        "clocks": primary["clocks"],
        "configs0": configs0,
    }
    if general is not None:
        line["general_input"] = general
    if secondary:
        line["secondary"] = secondary
    if measurement is not None:
        line["measurement"] = measurement
    emit(json.dumps(line))


class JsonStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner from C): the process's file descriptor 1 is pointed at stderr for the whole run and the
    line is written to the saved descriptor at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.fd, (line + "\n").encode())


OUT = None


def emit(line):
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line)


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="", choices=[""] + sorted(DEFAULTS),
                    help="default: qft (33 qubits, complex128) on one GPU, supremacy (34 qubits, complex64) on several")
    ap.add_argument("--nqubits", type=int, default=0)
    ap.add_argument("--dtype", default="")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads of the default run")
    ap.add_argument("--no-n1", action="store_true",
                    help="--gpus N > 1: skip the single-GPU run of the same circuit on rank 0 (the N = 1 point)")
    ap.add_argument("--measure", action="store_true",
                    help="append the measurement leg (probabilities, 10^6 shots, collapse); default for --workload qv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    OUT = JsonStdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
