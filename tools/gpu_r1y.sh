#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -40
timeout 600 python -m pytest tests/test_program_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/prog_bench.py --workload qft --nqubits 33 2>&1 | tail -3
