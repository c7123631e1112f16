set -x
ncu --set full --clock-control none --import-source on -k regex:k_tile_program -c 1 -o gpurun_out/r1_prof_tileprog python tools/prog_bench.py --workload variational --nqubits 28 --reps 1 > gpurun_out/r1_prof_tileprog.log 2>&1
tail -3 gpurun_out/r1_prof_tileprog.log
