#!/usr/bin/env python
"""Add to tests/golden/marginals.json the 4-qubit marginal of a benchmark circuit that is too large
for any host (supremacy-34 complex64: 137 GB), from a 1-GPU run of this framework:

    python tools/gpu_marginal_fixture.py supremacy-34-complex64 [more keys]     (on a GPU box)

The entry says where it comes from.  It pins the MULTI-GPU runs (bench.py --gpus N compares its
sharded marginal with it); the single-GPU path that produced it is itself pinned against the
reference's numba kernels by the smaller entries of the same circuit family (supremacy-30, made by
tests/golden/make_marginals.py), which this script re-checks before writing anything."""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from qibojit_b200.backends.b200 import B200Backend

    path = os.path.join(ROOT, "tests", "golden", "marginals.json")
    store = json.load(open(path))
    b = B200Backend()
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)

    def marginal(workload, n, dtype):
        b.set_dtype(dtype)
        state = b.execute_circuit(bench.build_circuit(workload, n))
        p = b.calculate_probabilities(state, bench.MARGINAL_QUBITS, n).double().cpu().numpy()
        del state
        torch.cuda.empty_cache()
        return p

    for key in sys.argv[1:]:
        workload, n, dtype = key.split("-")
        n = int(n)
        # chain of trust: the same code path against the reference-made entry of the family
        small = next((k for k in store if k.startswith(workload + "-") and k.endswith(dtype)
                      and "reference numba" in store[k]["source"]), None)
        assert small is not None, f"no reference-made marginal of the {workload} family"
        sw, sn, sd = small.split("-")
        err = float(np.abs(marginal(sw, int(sn), sd) - np.asarray(store[small]["marginal"])).max())
        tol = bench.TOL[dtype]
        print(f"{small}: max |marginal - reference marginal| = {err:.3e} (tolerance {tol})")
        assert err <= tol
        p = marginal(workload, n, dtype)
        store[key] = {"marginal": [float(x) for x in p], "qubits": bench.MARGINAL_QUBITS,
                      "source": f"qibojit_b200 single-GPU run (tools/gpu_marginal_fixture.py); the same path matches the "
                                f"reference-made {small} entry to {err:.1e}"}
        print(key, "sum", float(p.sum()))
    with open(os.path.join(out_dir, "marginals.json"), "w") as f:
        json.dump(store, f, indent=1, sort_keys=True)
    print("wrote gpurun_out/marginals.json (copy it to tests/golden/marginals.json)")


if __name__ == "__main__":
    main()
