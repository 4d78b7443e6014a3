#!/bin/bash
# round 2, call A: pipe rates + FB-mask kernel variants (bit-exactness via mask sha)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_a_smi.txt
profiles/mb/bin/pipe_rates > gpurun_out/r02_a_pipe_rates.txt 2>&1
for v in 0 1 2 3 4 5 6 7 8; do
  PIXPRO_B200_FBTILE=$v timeout 120 python profiles/mb/fb_variants.py 64 1 >> gpurun_out/r02_a_fb_variants.txt 2>&1
done
cat gpurun_out/r02_a_pipe_rates.txt gpurun_out/r02_a_fb_variants.txt
