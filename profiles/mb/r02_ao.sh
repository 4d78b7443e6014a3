#!/bin/bash
mkdir -p gpurun_out
timeout 120 python profiles/mb/conv7_bench.py > gpurun_out/r02_ao_plain.log 2>&1; tail -8 gpurun_out/r02_ao_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 15 -c 3 -o gpurun_out/r02_ao_conv7 python profiles/mb/conv7_bench.py > gpurun_out/r02_ao_ncu.log 2>&1; echo "ncu rc=$?"
