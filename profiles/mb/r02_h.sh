#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_corr.py -q 2>&1 | tail -15 > gpurun_out/r02_h_corr.log; cat gpurun_out/r02_h_corr.log
timeout 300 python -m pytest tests/test_gpu_tc.py -q -k "tc_gemm" 2>&1 | tail -25 > gpurun_out/r02_h_tc.log; cat gpurun_out/r02_h_tc.log
timeout 300 python profiles/mb/tc_gemm_bench.py > gpurun_out/r02_h_tcbench.txt 2>&1; cat gpurun_out/r02_h_tcbench.txt
