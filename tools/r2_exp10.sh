#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2k
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest.log
timeout 1200 python bench.py --steps 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -c 300 ${O}_bench.json; tail -3 ${O}_bench.err
