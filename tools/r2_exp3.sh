#!/bin/bash
# round-2 experiment 3: two-stage F factors, in-place 4x4 bodies; new bench.py end to end
mkdir -p gpurun_out
O=gpurun_out/r2c
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest.log
PB="timeout 600 python tools/prog_bench.py --reps 3"
{
$PB --workload variational --nqubits 30 --run-bits 3
$PB --workload qft --nqubits 30 --run-bits 3
QJ_DIAGF_MIN=0 $PB --workload qft --nqubits 30 --run-bits 3
$PB --workload supremacy --nqubits 32 --dtype complex64
$PB --workload qft --nqubits 33 --run-bits 3
$PB --workload qft --nqubits 33 --run-bits 3 --keep-swaps
} > ${O}_prog_bench.txt 2>&1
cat ${O}_prog_bench.txt | grep -v "^  pass"
timeout 900 python bench.py --steps 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -c 1500 ${O}_bench.json; tail -5 ${O}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_var30 python tools/prog_bench.py --workload variational --nqubits 30 --run-bits 3 --reps 1 > ${O}_ncu_var30.log 2>&1; echo "ncu var rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_qft30 python tools/prog_bench.py --workload qft --nqubits 30 --run-bits 3 --reps 1 > ${O}_ncu_qft30.log 2>&1; echo "ncu qft rc=$?"
cp qibojit_b200/lib/libqibojit_b200.so ${O}_lib.so
ls -la gpurun_out/ | tail -12
