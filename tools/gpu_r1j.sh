#!/bin/bash
# multi-GPU checks: distributed parity (NCCL), exchange bandwidth, bench legs
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1j}
N=${NGPU:-2}
G=$(python -c "print(int($N).bit_length()-1)")
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_dist_check.log | tail -24
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_variational30_g$N.json 2> gpurun_out/${TAG}_bench_variational30_g$N.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_variational30_g$N.json; tail -3 gpurun_out/${TAG}_bench_variational30_g$N.err
NQ=$((33 + G))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --workload supremacy --nqubits $NQ > gpurun_out/${TAG}_bench_supremacy${NQ}_g$N.json 2> gpurun_out/${TAG}_bench_supremacy${NQ}_g$N.err; echo "sup rc=$?"; cat gpurun_out/${TAG}_bench_supremacy${NQ}_g$N.json; tail -3 gpurun_out/${TAG}_bench_supremacy${NQ}_g$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --workload qft --nqubits $NQ > gpurun_out/${TAG}_bench_qft${NQ}_g$N.json 2> gpurun_out/${TAG}_bench_qft${NQ}_g$N.err; echo "qft rc=$?"; cat gpurun_out/${TAG}_bench_qft${NQ}_g$N.json; tail -3 gpurun_out/${TAG}_bench_qft${NQ}_g$N.err
if [ -n "$SINGLE" ]; then
for args in "--workload variational --nqubits 30" "--workload qft --nqubits 33" "--workload supremacy --nqubits 32 --dtype complex64"; do
  timeout 300 python tools/prog_bench.py $args --out gpurun_out/${TAG}_prog_bench.jsonl 2>&1 | tail -16
done
fi
ls -la gpurun_out
