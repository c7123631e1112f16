#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1u}
timeout 1200 python -m pytest tests/test_program_gpu.py -x -q 2>&1 | tail -3
for args in "--workload supremacy --nqubits 32 --dtype complex64" "--workload qv --nqubits 30 --dtype complex64" "--workload qft --nqubits 31 --dtype complex64" "--workload variational --nqubits 31 --dtype complex64"; do
  timeout 300 python tools/prog_bench.py $args --out gpurun_out/${TAG}_prog_bench.jsonl 2>&1 | tail -18
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/${TAG}_sup32_pass python tools/prog_bench.py --workload supremacy --nqubits 32 --dtype complex64 --reps 1 > gpurun_out/${TAG}_ncu_sup.log 2>&1
