#!/bin/bash
mkdir -p gpurun_out
for v in 1 3 4 9 10; do PIXPRO_B200_FBTILE=$v timeout 120 python profiles/mb/fb_variants.py 64 1; done > gpurun_out/r02_g_fb.txt 2>&1
cat gpurun_out/r02_g_fb.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python main_pretrain.py --synthetic --batch-size 64 --n-frames 6 --steps 10 --warmup 5 --print-freq 5 > gpurun_out/r02_g_pretrain1.json 2> gpurun_out/r02_g_pretrain1.err; tail -3 gpurun_out/r02_g_pretrain1.err; cat gpurun_out/r02_g_pretrain1.json
timeout 600 python bench.py --pretrain --pretrain-batch 128 --steps 8 --warmup 4 > gpurun_out/r02_g_bench_pretrain.json 2> gpurun_out/r02_g_bench_pretrain.err; tail -3 gpurun_out/r02_g_bench_pretrain.err; cat gpurun_out/r02_g_bench_pretrain.json
