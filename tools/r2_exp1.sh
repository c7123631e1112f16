#!/bin/bash
# round-2 experiment 1: parity of the new ops on the GPU, then tile geometry / fusion sweeps + ncu
mkdir -p gpurun_out
O=gpurun_out/r2a
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
PB="timeout 600 python tools/prog_bench.py --reps 3"
{
for cfg in "11 4" "11 3" "12 4" "12 3"; do set -- $cfg
  $PB --workload variational --nqubits 30 --tile-bits $1 --run-bits $2
done
$PB --workload variational --nqubits 30 --no-pair-blocks
for m in 0 2 3; do
  QJ_DIAGF_MIN=$m $PB --workload qft --nqubits 30
  QJ_DIAGF_MIN=$m $PB --workload qft --nqubits 30 --tile-bits 12
done
$PB --workload qft --nqubits 30 --tile-bits 12 --run-bits 3
$PB --workload qft --nqubits 30 --keep-swaps
$PB --workload supremacy --nqubits 32 --dtype complex64
$PB --workload supremacy --nqubits 32 --dtype complex64 --no-pair-blocks
QJ_DIAGF_MIN=0 $PB --workload supremacy --nqubits 32 --dtype complex64
$PB --workload qv --nqubits 30 --dtype complex64
$PB --workload qft --nqubits 33
$PB --workload qft --nqubits 33 --tile-bits 12
$PB --workload qft --nqubits 33 --keep-swaps
} > ${O}_prog_bench.txt 2>&1
cat ${O}_prog_bench.txt | grep -v "^  pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_var30 python tools/prog_bench.py --workload variational --nqubits 30 --reps 1 > ${O}_ncu_var30.log 2>&1; echo "ncu var rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_qft30 python tools/prog_bench.py --workload qft --nqubits 30 --reps 1 > ${O}_ncu_qft30.log 2>&1; echo "ncu qft rc=$?"
ls -la gpurun_out/
