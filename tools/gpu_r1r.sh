#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1r}
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --workload supremacy --nqubits 36 > gpurun_out/${TAG}_bench_supremacy36_g$N.json 2> gpurun_out/${TAG}_bench_supremacy36_g$N.err; echo "sup rc=$?"; cat gpurun_out/${TAG}_bench_supremacy36_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_supremacy36_g$N.err | tail -3
