#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2j
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
PB="timeout 600 python tools/prog_bench.py --reps 3"
{
$PB --workload variational --nqubits 30
$PB --workload qft --nqubits 30
$PB --workload supremacy --nqubits 32 --dtype complex64
$PB --workload supremacy --nqubits 34 --dtype complex64
$PB --workload qv --nqubits 30 --dtype complex64
} > ${O}_prog_bench.txt 2>&1
cat ${O}_prog_bench.txt | grep -v "^  pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_sup30 python tools/prog_bench.py --workload supremacy --nqubits 30 --dtype complex64 --reps 1 > ${O}_ncu_sup30.log 2>&1; echo "ncu sup rc=$?"
cp qibojit_b200/lib/libqibojit_b200.so ${O}_lib.so
