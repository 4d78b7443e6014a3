#!/bin/bash
# historical: PIXPRO_B200_CONVPAD selected the padded TMA route of the 7x7 value transform, removed after these measurements (profiles/r02_w_convpad_step.txt)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -15 > gpurun_out/r02_v_tests.log; cat gpurun_out/r02_v_tests.log
{ timeout 120 python profiles/mb/conv7_bench.py; PIXPRO_B200_CONVPAD=0 timeout 120 python profiles/mb/conv7_bench.py; } > gpurun_out/r02_v_conv7.txt 2>&1
cat gpurun_out/r02_v_conv7.txt
