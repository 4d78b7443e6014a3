#!/bin/bash
mkdir -p gpurun_out
for ns in 0 3000 6000 10000; do
  echo "== stagger $ns ns"
  PIXPRO_B200_TC2_STAGGER=$ns timeout 300 python profiles/mb/tc_gemm_bench.py 2>&1 | grep "TMA-fed" | cut -c1-140
  PIXPRO_B200_TC2_STAGGER=$ns timeout 600 python bench.py --batch 32 --grid 28 --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('   g28 step %.3f ms; S %.3f gS %.3f Y %.3f gvh %.3f gxh %.3f loss %.3f' % (d['ms_per_step'], k['ppm S (tcgen05)']['ms_per_step'], k['ppm gS (tcgen05)']['ms_per_step'], k['ppm Y (tcgen05)']['ms_per_step'], k['ppm gvh (tcgen05)']['ms_per_step'], k['ppm gxh (tcgen05)']['ms_per_step'], k['loss M=K*pos^T (tcgen05)']['ms_per_step']))"
done > gpurun_out/r02_s_stagger.txt 2>&1
cat gpurun_out/r02_s_stagger.txt
