#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1v}
timeout 1200 python -m pytest tests/test_program_gpu.py -x -q 2>&1 | tail -3
for args in "--workload qft --nqubits 33" "--workload variational --nqubits 30" "--workload supremacy --nqubits 32 --dtype complex64" "--workload qft --nqubits 31 --dtype complex64"; do
  timeout 300 python tools/prog_bench.py $args --out gpurun_out/${TAG}_prog_bench.jsonl 2>&1 | tail -18
done
