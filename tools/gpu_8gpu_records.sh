#!/bin/bash
# 8-GPU record runs of the named sharded workloads
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1r}
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --workload supremacy --nqubits 36 > gpurun_out/${TAG}_bench_supremacy36_g$N.json 2> gpurun_out/${TAG}_bench_supremacy36_g$N.err; echo "sup rc=$?"; cat gpurun_out/${TAG}_bench_supremacy36_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_supremacy36_g$N.err | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 2 --warmup 3 --workload qft --nqubits 36 > gpurun_out/${TAG}_bench_qft36c128_g$N.json 2> gpurun_out/${TAG}_bench_qft36c128_g$N.err; echo "qft rc=$?"; cat gpurun_out/${TAG}_bench_qft36c128_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_qft36c128_g$N.err | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_variational30_g$N.json 2> gpurun_out/${TAG}_bench_variational30_g$N.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_variational30_g$N.json
