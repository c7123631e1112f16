#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1z}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_variational30.json 2> gpurun_out/${TAG}_bench_variational30.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_variational30.json; tail -5 gpurun_out/${TAG}_bench_variational30.err
timeout 900 python bench.py --workload qv --steps 2 > gpurun_out/${TAG}_bench_qv32.json 2> gpurun_out/${TAG}_bench_qv32.err; echo "qv rc=$?"; cat gpurun_out/${TAG}_bench_qv32.json; tail -5 gpurun_out/${TAG}_bench_qv32.err
timeout 900 python bench.py --workload qft --steps 3 > gpurun_out/${TAG}_bench_qft33.json 2> gpurun_out/${TAG}_bench_qft33.err; echo "qft rc=$?"; cat gpurun_out/${TAG}_bench_qft33.json; tail -5 gpurun_out/${TAG}_bench_qft33.err
timeout 900 python bench.py --workload supremacy --steps 3 > gpurun_out/${TAG}_bench_supremacy32.json 2> gpurun_out/${TAG}_bench_supremacy32.err; echo "sup rc=$?"; cat gpurun_out/${TAG}_bench_supremacy32.json; tail -5 gpurun_out/${TAG}_bench_supremacy32.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_variational30.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 0.5 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python -c "import __graft_entry__ as g; g.smoke()"
