set -x
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r1_pytest_gpu.log
python bench.py > gpurun_out/r1_bench_variational30.json 2> gpurun_out/r1_bench_variational30.err; echo "bench rc=$?"; cat gpurun_out/r1_bench_variational30.json; tail -3 gpurun_out/r1_bench_variational30.err
python tools/sweep.py --nqubits 30 --dtype complex128 --special --out gpurun_out/r1_sweep_c128_n30.json > gpurun_out/r1_sweep_c128_n30.log 2>&1; echo "sweep rc=$?"
python tools/sweep.py --nqubits 31 --dtype complex64 --special --out gpurun_out/r1_sweep_c64_n31.json > gpurun_out/r1_sweep_c64_n31.log 2>&1; echo "sweep rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_variational30.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 0.5 > gpurun_out/r1_ncu_bench.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_dense_direct -s 20 -c 3 -o gpurun_out/r1_prof_dense python bench.py --steps 1 --warmup 3 --cpu-seconds 0.5 > gpurun_out/r1_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
