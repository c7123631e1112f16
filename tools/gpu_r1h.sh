#!/bin/bash
# pass kernel v4 (uniform op loop, one-target group ops, runtime CTA size): parity, timing, ncu
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1m}
timeout 1200 python -m pytest tests/test_program_gpu.py -x -q 2>&1 | tail -5
for args in "--workload variational --nqubits 30" "--workload variational --nqubits 30 --tile-bits 11" \
            "--workload qft --nqubits 30" "--workload qft --nqubits 30 --tile-bits 11" "--workload qft --nqubits 33" \
            "--workload supremacy --nqubits 32 --dtype complex64" "--workload supremacy --nqubits 32 --dtype complex64 --tile-bits 12" \
            "--workload qv --nqubits 30"; do
  timeout 300 python tools/prog_bench.py $args --out gpurun_out/${TAG}_prog_bench.jsonl 2>&1 | tail -16
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/${TAG}_var30_pass python tools/prog_bench.py --workload variational --nqubits 30 --reps 1 > gpurun_out/${TAG}_ncu_var.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/${TAG}_qft30_pass python tools/prog_bench.py --workload qft --nqubits 30 --reps 1 > gpurun_out/${TAG}_ncu_qft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 1 -o gpurun_out/${TAG}_sup32_pass python tools/prog_bench.py --workload supremacy --nqubits 32 --dtype complex64 --reps 1 > gpurun_out/${TAG}_ncu_sup.log 2>&1
ls -la gpurun_out
