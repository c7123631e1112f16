#!/bin/bash
# round-2 experiment 6: evidence run -- bench.py end to end, ncu launch list of the bench command,
# DRAM traffic of the QFT-33 passes, full ncu of the variational pass, DMMA vs DFMA microbenchmark
mkdir -p gpurun_out
O=gpurun_out/r2f
timeout 1200 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -c 600 ${O}_bench.json; tail -3 ${O}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-secondary --cpu-seconds 1 > ${O}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --replay-mode application --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_pass -c 4 -o ${O}_ncu_qft33_traffic python tools/prog_bench.py --workload qft --nqubits 33 --reps 0 > ${O}_ncu_qft33.log 2>&1; echo "ncu qft33 traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_var30 python tools/prog_bench.py --workload variational --nqubits 30 --reps 1 > ${O}_ncu_var30.log 2>&1; echo "ncu var rc=$?"
tools/micro/dmma_bench > ${O}_dmma_bench.json; cat ${O}_dmma_bench.json
timeout 300 ncu --clock-control none --metrics sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_tensor_op_dmma.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --csv --log-file ${O}_dmma_ncu.csv tools/micro/dmma_bench > /dev/null 2>&1; echo "ncu dmma rc=$?"; tail -20 ${O}_dmma_ncu.csv | cut -c1-200
cp qibojit_b200/lib/libqibojit_b200.so ${O}_lib.so
