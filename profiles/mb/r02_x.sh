#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_x_tests.log; cat gpurun_out/r02_x_tests.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/r02_x_bench.json 2> gpurun_out/r02_x_bench.err; echo "bench rc=$? wall=${SECONDS}s"
SECONDS=0
timeout 600 python bench.py --impl reference > gpurun_out/r02_x_ref.json 2> gpurun_out/r02_x_ref.err; echo "ref rc=$? wall=${SECONDS}s"
tail -c 600 gpurun_out/r02_x_ref.json
