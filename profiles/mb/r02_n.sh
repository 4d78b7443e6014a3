#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_syncbn.py -x -q 2>&1 | tail -15 > gpurun_out/r02_n_tests.log; cat gpurun_out/r02_n_tests.log
grep -q " passed" gpurun_out/r02_n_tests.log && ! grep -q "failed" gpurun_out/r02_n_tests.log || exit 1
for mode in "" "--no-fast-syncbn"; do
  timeout 600 python main_pretrain.py --synthetic --batch-size 128 --n-frames 6 --steps 10 --warmup 5 --print-freq 100 $mode 2>/dev/null | tail -1 | cut -c1-400
done
