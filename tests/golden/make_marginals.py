"""Generate tests/golden/marginals.json: 4-qubit marginals of the benchmark circuits at (or near)
benchmark size, computed gate by gate with the REFERENCE's own numba kernels (oracle/_ref, the
unmodified custom_operators/gates.py of qibojit) through the dispatch restated in
tests/refdispatch.py.  bench.py and the -m gpu tests compare the CUDA path's marginals with these.

Run in the build container (needs oracle/_ref, i.e. /root/reference, numba and ~20 GB of RAM):

    python tests/golden/make_marginals.py [workload-n-dtype ...]

Entries whose state does not fit a host (supremacy-34 complex64: 137 GB) cannot be made here; they
are added by tools/gpu_marginal_fixture.py from a 1-GPU run of this framework, which is itself
pinned by the smaller entries of the same circuit family made here (same file, "source" says which).
"""

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

DEFAULT = ["variational-30-complex128", "supremacy-30-complex64", "qv-26-complex64", "variational-26-complex128"]
PATH = os.path.join(HERE, "marginals.json")


def main():
    wanted = sys.argv[1:] or DEFAULT
    try:
        with open(PATH) as f:
            store = json.load(f)
    except Exception:
        store = {}
    G, O, threads, kind = bench.cpu_kernels()
    assert kind == "reference", "oracle/_ref is missing: run python oracle/make_ref.py"
    for key in wanted:
        workload, n, dtype = key.split("-")
        n = int(n)
        circuit = bench.build_circuit(workload, n)
        prog = bench.reference_program(circuit, G)
        st = np.empty(1 << n, dtype=dtype)
        O.initial_state_vector(st)
        t0 = time.perf_counter()
        for call in prog:
            call(st, dtype)
        dt = time.perf_counter() - t0
        p = np.zeros(16)
        chunk = st.reshape(16, -1)
        for i in range(16):
            p[i] = float(np.sum(chunk[i].real.astype(np.float64) ** 2) + np.sum(chunk[i].imag.astype(np.float64) ** 2))
        store[key] = {"marginal": [float(x) for x in p], "qubits": bench.MARGINAL_QUBITS,
                      "source": f"reference numba kernels (oracle/_ref), gate by gate, {threads} threads "
                                f"(tests/golden/make_marginals.py)",
                      "seconds": round(dt, 1)}
        print(key, "sum", p.sum(), f"{dt:.1f}s", flush=True)
        del st, chunk
        with open(PATH, "w") as f:
            json.dump(store, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
