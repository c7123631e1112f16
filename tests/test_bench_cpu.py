"""CPU checks of bench.py's reference arm (the GPU arm needs a device): one JSON line on stdout
with the keys of the bench contract."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nqubits", "16",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gates_per_second" and d["unit"] == "gates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["config"]["workload"] == "qft-16-complex128"
    # the reference's own numba kernels when oracle/_ref is there (the build container), else the C port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "qibojit", "custom_operators", "gates.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ms_per_step is the time of one whole circuit at the measured rate, not the wall time of a sample
    assert abs(d["ms_per_step"] - 1e3 * d["config"]["circuit_gates"] / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["configs0"]["workload"] == "qft-20-complex128" and d["configs0"]["max_abs_err_vs_closed_form"] < 1e-12


def test_reference_arm_ignores_omp_num_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference sizes its thread pool from the CPU affinity
    (backends/cpu.py:86-89) and so must the CPU arm."""
    import psutil

    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--nqubits", "16", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["cores"] == len(psutil.Process().cpu_affinity())
    assert d["config"]["workload"] == "supremacy-16-complex64"


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--nqubits", "16", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
