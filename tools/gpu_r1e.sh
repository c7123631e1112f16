#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_program_gpu.py -x -q 2>&1 | tail -25
ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/r1e_qft30_pass0 python tools/prog_bench.py --workload qft --nqubits 30 --reps 1 > gpurun_out/r1e_ncu_qft.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pass --launch-skip 2 -c 1 -o gpurun_out/r1e_qv30_pass2 python tools/prog_bench.py --workload qv --nqubits 30 --reps 1 > gpurun_out/r1e_ncu_qv.log 2>&1
tail -3 gpurun_out/r1e_ncu_qft.log gpurun_out/r1e_ncu_qv.log
