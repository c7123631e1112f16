#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1w}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/${TAG}_qft30_pass python tools/prog_bench.py --workload qft --nqubits 30 --reps 1 > gpurun_out/${TAG}_ncu_qft.log 2>&1
