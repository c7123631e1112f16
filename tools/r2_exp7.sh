#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
timeout 900 python tools/gpu_marginal_fixture.py supremacy-34-complex64 supremacy-32-complex64 2>&1 | tail -6
timeout 900 ncu --replay-mode application --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_pass -c 4 -o ${O}_ncu_qft33_traffic python tools/prog_bench.py --workload qft --nqubits 33 --reps 1 > ${O}_ncu_qft33.log 2>&1; echo "ncu qft33 traffic rc=$?"; tail -3 ${O}_ncu_qft33.log
