set -x
python -m pytest tests/test_program_gpu.py -x -q 2>&1 | tail -15
for args in "--workload variational --nqubits 30" "--workload qft --nqubits 30" "--workload qft --nqubits 33" "--workload supremacy --nqubits 32 --dtype complex64" "--workload qv --nqubits 32 --dtype complex64"; do
  timeout 300 python tools/prog_bench.py $args --out gpurun_out/r1_prog_bench4.jsonl 2>&1 | tail -14
done
ncu --set full --clock-control none --import-source on -k regex:k_tile_program -c 1 -o gpurun_out/r1_prof_tileprog2 python tools/prog_bench.py --workload variational --nqubits 28 --reps 1 > gpurun_out/r1_prof_tileprog2.log 2>&1
