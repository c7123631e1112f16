#!/bin/bash
# final single-GPU evidence of the round: smoke, GPU test suite, both bench arms, ncu launch list of the
# bench command, full ncu capture of the pass kernel (QFT-30: the benchmark's kernel at a size ncu replays quickly)
mkdir -p gpurun_out
O=gpurun_out/r2final
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${O}_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
timeout 1200 python bench.py --steps 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -c 400 ${O}_bench.json; tail -3 ${O}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "reference rc=$?"; tail -c 600 ${O}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-secondary --cpu-seconds 1 > ${O}_launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 0 -c 2 -o ${O}_ncu_qft30 python tools/prog_bench.py --workload qft --nqubits 30 --reps 1 > ${O}_ncu_qft30.log 2>&1; echo "ncu qft rc=$?"
cp qibojit_b200/lib/libqibojit_b200.so ${O}_lib.so
