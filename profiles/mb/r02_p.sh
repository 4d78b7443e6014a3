#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_corr.py -x -q 2>&1 | tail -25 > gpurun_out/r02_p_tests.log; cat gpurun_out/r02_p_tests.log
