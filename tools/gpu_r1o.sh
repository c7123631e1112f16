#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1o}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/${TAG}_qft30_pass python tools/prog_bench.py --workload qft --nqubits 30 --reps 1 > gpurun_out/${TAG}_ncu_qft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/${TAG}_var30_pass python tools/prog_bench.py --workload variational --nqubits 30 --reps 1 > gpurun_out/${TAG}_ncu_var.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/${TAG}_sup32_pass python tools/prog_bench.py --workload supremacy --nqubits 32 --dtype complex64 --reps 1 > gpurun_out/${TAG}_ncu_sup.log 2>&1
for args in "--workload supremacy --nqubits 32 --dtype complex64 --tile-bits 12"; do
  timeout 300 python tools/prog_bench.py $args 2>&1 | tail -18
done
