#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2l
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --workload qv --steps 2 --no-secondary --cpu-seconds 2 > ${O}_bench_qv32.json 2> ${O}_bench_qv32.err; echo "qv rc=$?"; python -c "
import json; d=json.load(open('${O}_bench_qv32.json')); print(d['value'], d['ms_per_step'], d['measurement'])"; tail -2 ${O}_bench_qv32.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_sup30 python tools/prog_bench.py --workload supremacy --nqubits 30 --dtype complex64 --reps 1 > ${O}_ncu_sup30.log 2>&1; echo "ncu sup rc=$?"
cp qibojit_b200/lib/libqibojit_b200.so ${O}_lib.so
