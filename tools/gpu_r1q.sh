#!/bin/bash
# 8-GPU run: distributed parity, variational-30 (strong scaling point), supremacy-36 and QFT-36 complex64
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1q}
N=8
QJ_NLOCAL=28 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_dist_check.log | tail -22
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_variational30_g$N.json 2> gpurun_out/${TAG}_bench_variational30_g$N.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_variational30_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_variational30_g$N.err | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --workload supremacy --nqubits 36 > gpurun_out/${TAG}_bench_supremacy36_g$N.json 2> gpurun_out/${TAG}_bench_supremacy36_g$N.err; echo "sup rc=$?"; cat gpurun_out/${TAG}_bench_supremacy36_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_supremacy36_g$N.err | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 2 --warmup 3 --workload qft --nqubits 36 --dtype complex64 > gpurun_out/${TAG}_bench_qft36c64_g$N.json 2> gpurun_out/${TAG}_bench_qft36c64_g$N.err; echo "qft rc=$?"; cat gpurun_out/${TAG}_bench_qft36c64_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_qft36c64_g$N.err | tail -3
ls -la gpurun_out
