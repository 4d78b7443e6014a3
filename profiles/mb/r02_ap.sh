#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -k "conv or featprop" 2>&1 | tail -3
timeout 120 python profiles/mb/conv7_bench.py 2>&1 | tail -8
