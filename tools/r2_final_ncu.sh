#!/bin/bash
# full ncu captures of the final pass kernel on the complex64 and the variational workloads
mkdir -p gpurun_out
O=gpurun_out/r2final
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_sup30 python tools/prog_bench.py --workload supremacy --nqubits 30 --dtype complex64 --reps 1 > ${O}_ncu_sup30.log 2>&1; echo "ncu sup rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_var30 python tools/prog_bench.py --workload variational --nqubits 30 --reps 1 > ${O}_ncu_var30.log 2>&1; echo "ncu var rc=$?"
cp qibojit_b200/lib/libqibojit_b200.so ${O}_lib.so
