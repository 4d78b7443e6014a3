#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -k "tc_gemm" 2>&1 | tail -25 > gpurun_out/r02_i_tc.log; cat gpurun_out/r02_i_tc.log
grep -q " passed" gpurun_out/r02_i_tc.log && ! grep -q "failed" gpurun_out/r02_i_tc.log || exit 1
timeout 300 python profiles/mb/tc_gemm_bench.py > gpurun_out/r02_i_tcbench.txt 2>&1; cat gpurun_out/r02_i_tcbench.txt
PIXPRO_B200_TC2=2 timeout 300 python profiles/mb/tc_gemm_bench.py > gpurun_out/r02_i_tcbench_mode2.txt 2>&1; cat gpurun_out/r02_i_tcbench_mode2.txt
timeout 120 python profiles/run_tc_gemm.py > gpurun_out/r02_i_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc2_gemm -c 1 -o gpurun_out/r02_i_tc2 python profiles/run_tc_gemm.py > gpurun_out/r02_i_ncu.log 2>&1
tail -3 gpurun_out/r02_i_ncu.log
