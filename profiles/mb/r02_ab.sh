#!/bin/bash
mkdir -p gpurun_out
for v in 1 3 4; do PIXPRO_B200_FBUP=0 PIXPRO_B200_FBTILE=$v timeout 120 python profiles/mb/fb_variants.py 64 1; done > gpurun_out/r02_ab_fb_boxrows.txt 2>&1
cat gpurun_out/r02_ab_fb_boxrows.txt
