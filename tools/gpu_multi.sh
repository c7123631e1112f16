#!/bin/bash
# multi-GPU checks (NGPU=2/4/8): pack kernels, distributed parity on NCCL, exchange bandwidth, bench legs
set -x
mkdir -p gpurun_out
TAG=${TAG:-multi}
N=${NGPU:-2}
G=$(python -c "print(int($N).bit_length()-1)")
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k swap_pack 2>&1 | tail -3
QJ_NLOCAL=${QJ_NLOCAL:-28} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check_g$N.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_dist_check_g$N.log | tail -24
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_variational30_g$N.json 2> gpurun_out/${TAG}_bench_variational30_g$N.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_variational30_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_variational30_g$N.err | tail -3
NQ=$((33 + G))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --workload supremacy --nqubits $NQ > gpurun_out/${TAG}_bench_supremacy${NQ}_g$N.json 2> gpurun_out/${TAG}_bench_supremacy${NQ}_g$N.err; echo "sup rc=$?"; cat gpurun_out/${TAG}_bench_supremacy${NQ}_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_supremacy${NQ}_g$N.err | tail -3
if [ -n "$QFT" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 2 --warmup 3 --workload qft --nqubits $NQ > gpurun_out/${TAG}_bench_qft${NQ}_g$N.json 2> gpurun_out/${TAG}_bench_qft${NQ}_g$N.err; echo "qft rc=$?"; cat gpurun_out/${TAG}_bench_qft${NQ}_g$N.json; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_bench_qft${NQ}_g$N.err | tail -3
fi
