#!/bin/bash
# round-2 experiment 2: program image in the constant bank, constant / fused diagonals, 128-byte runs
mkdir -p gpurun_out
O=gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
PB="timeout 600 python tools/prog_bench.py --reps 3"
{
$PB --workload variational --nqubits 30 --run-bits 4
$PB --workload variational --nqubits 30 --run-bits 3
$PB --workload variational --nqubits 30 --run-bits 2
for m in 0 2 3; do
  QJ_DIAGF_MIN=$m $PB --workload qft --nqubits 30 --run-bits 3
done
$PB --workload qft --nqubits 30 --run-bits 4
$PB --workload qft --nqubits 30 --run-bits 3 --keep-swaps
$PB --workload supremacy --nqubits 32 --dtype complex64 --run-bits 5
$PB --workload supremacy --nqubits 32 --dtype complex64 --run-bits 4
QJ_DIAGF_MIN=0 $PB --workload supremacy --nqubits 32 --dtype complex64 --run-bits 4
$PB --workload qv --nqubits 30 --dtype complex64
$PB --workload qft --nqubits 33 --run-bits 3
$PB --workload qft --nqubits 33 --run-bits 3 --keep-swaps
QJ_DIAGF_MIN=0 $PB --workload qft --nqubits 33 --run-bits 3
} > ${O}_prog_bench.txt 2>&1
cat ${O}_prog_bench.txt | grep -v "^  pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_var30 python tools/prog_bench.py --workload variational --nqubits 30 --run-bits 3 --reps 1 > ${O}_ncu_var30.log 2>&1; echo "ncu var rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_qft30 python tools/prog_bench.py --workload qft --nqubits 30 --run-bits 3 --reps 1 > ${O}_ncu_qft30.log 2>&1; echo "ncu qft rc=$?"
cp qibojit_b200/lib/libqibojit_b200.so ${O}_lib.so
ls -la gpurun_out/
