#!/bin/bash
# multi-GPU checks (NGPU=2/4/8): distributed parity on both transports, measurement, exchange bandwidth, bench legs
mkdir -p gpurun_out
TAG=${TAG:-r2m}
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
filter() { grep -v "^W\|^\[W\|^\*\|^$\|OMP_NUM\|Gloo\|NCCL WARN"; }
[ -z "$SKIP_CHECK" ] && QJ_NLOCAL=${QJ_NLOCAL:-29} timeout 900 $TR --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check_g$N.log 2>&1; echo "dist_check rc=$?"; [ -z "$SKIP_CHECK" ] && filter < gpurun_out/${TAG}_dist_check_g$N.log | tail -${TAILN:-60}
if [ -z "$SKIP_BENCH" ]; then
timeout 1500 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 > gpurun_out/${TAG}_bench_g$N.json 2> gpurun_out/${TAG}_bench_g$N.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench_g$N.json; filter < gpurun_out/${TAG}_bench_g$N.err | tail -5
[ -z "$SKIP_REF" ] && timeout 600 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference_g$N.json 2> gpurun_out/${TAG}_bench_reference_g$N.err; echo "reference rc=$?"; cat gpurun_out/${TAG}_bench_reference_g$N.json; filter < gpurun_out/${TAG}_bench_reference_g$N.err | tail -3
fi
