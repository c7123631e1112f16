#!/bin/bash
# round-1 GPU check of the v2 pass kernel: parity tests, then per-pass timing of the benchmark circuits
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_program_gpu.py -x -q 2>&1 | tail -25
TAG=${TAG:-r1d}
for args in "--workload variational --nqubits 30" "--workload qft --nqubits 30" "--workload qft --nqubits 33" \
            "--workload supremacy --nqubits 32 --dtype complex64" "--workload qv --nqubits 32 --dtype complex64" \
            "--workload qv --nqubits 30"; do
  timeout 300 python tools/prog_bench.py $args --out gpurun_out/${TAG}_prog_bench.jsonl 2>&1 | tail -16
done
