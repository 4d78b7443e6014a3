#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -v > gpurun_out/r02_k_tests.log 2>&1
grep -n "PASSED\|FAILED\|ERROR" gpurun_out/r02_k_tests.log | tail -5
grep -n "Error\|error\|illegal\|trap" gpurun_out/r02_k_tests.log | head -10
