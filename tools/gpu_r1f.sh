#!/bin/bash
# round-1 re-entry: full GPU parity suite, per-pass timing of the benchmark circuits, ncu of the pass kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r1f_pytest_gpu.log
for args in "--workload variational --nqubits 30" "--workload qft --nqubits 30" "--workload qft --nqubits 33" \
            "--workload supremacy --nqubits 32 --dtype complex64" "--workload qv --nqubits 32 --dtype complex64" \
            "--workload qv --nqubits 30"; do
  timeout 300 python tools/prog_bench.py $args --out gpurun_out/r1f_prog_bench.jsonl 2>&1 | tail -16
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 2 -o gpurun_out/r1f_var30_pass python tools/prog_bench.py --workload variational --nqubits 30 --reps 1 > gpurun_out/r1f_ncu_var.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 2 -o gpurun_out/r1f_sup32_pass python tools/prog_bench.py --workload supremacy --nqubits 32 --dtype complex64 --reps 1 > gpurun_out/r1f_ncu_sup.log 2>&1
tail -3 gpurun_out/r1f_ncu_var.log gpurun_out/r1f_ncu_sup.log
ls -la gpurun_out
