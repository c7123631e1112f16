set -x
python bench.py --workload qft --steps 2 --warmup 3 --cpu-seconds 10 > gpurun_out/r1_bench_qft33.json 2> gpurun_out/r1_bench_qft33.err; echo "rc=$?"; cat gpurun_out/r1_bench_qft33.json; tail -5 gpurun_out/r1_bench_qft33.err
python bench.py --workload supremacy --steps 3 --warmup 3 --cpu-seconds 10 > gpurun_out/r1_bench_supremacy32.json 2> gpurun_out/r1_bench_supremacy32.err; echo "rc=$?"; cat gpurun_out/r1_bench_supremacy32.json; tail -5 gpurun_out/r1_bench_supremacy32.err
python bench.py --workload qv --steps 2 --warmup 3 --cpu-seconds 10 > gpurun_out/r1_bench_qv32.json 2> gpurun_out/r1_bench_qv32.err; echo "rc=$?"; cat gpurun_out/r1_bench_qv32.json; tail -5 gpurun_out/r1_bench_qv32.err
