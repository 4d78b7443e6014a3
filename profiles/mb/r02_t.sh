#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 3 4 5 6; do PIXPRO_B200_FBTILE=$v timeout 120 python profiles/mb/fb_variants.py 64 1; done > gpurun_out/r02_t_fb_pitch.txt 2>&1
cat gpurun_out/r02_t_fb_pitch.txt
timeout 600 python -m pytest tests/test_gpu_model.py -q -k "flow_store" 2>&1 | tail -3
