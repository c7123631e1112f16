#!/bin/bash
# final multi-GPU evidence: default bench line (supremacy-34 primary) and the configs[4] leg (QV-32 + 10^6 shots + collapse)
mkdir -p gpurun_out
TAG=${TAG:-r2p}
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
filter() { grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM\|Gloo\|NCCL WARN"; }
timeout 1500 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 > gpurun_out/${TAG}_bench_g$N.json 2> gpurun_out/${TAG}_bench_g$N.err; echo "bench rc=$?"; wc -l gpurun_out/${TAG}_bench_g$N.json; tail -c 400 gpurun_out/${TAG}_bench_g$N.json; filter < gpurun_out/${TAG}_bench_g$N.err | tail -5
timeout 1500 $TR --master-port 29514 bench.py --gpus $N --workload qv --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_qv32_g$N.json 2> gpurun_out/${TAG}_bench_qv32_g$N.err; echo "qv rc=$?"; tail -c 1200 gpurun_out/${TAG}_bench_qv32_g$N.json; filter < gpurun_out/${TAG}_bench_qv32_g$N.err | tail -5
