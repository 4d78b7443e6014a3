#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/mb/corr_bench.py > gpurun_out/r02_q_corr_bench.txt 2>&1; cat gpurun_out/r02_q_corr_bench.txt
