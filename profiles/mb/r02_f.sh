#!/bin/bash
mkdir -p gpurun_out
for v in 1 2; do PIXPRO_B200_CHAINUP=$v python profiles/mb/chain_up_variants.py 64 5; done > gpurun_out/r02_f_chain.txt 2>&1
cat gpurun_out/r02_f_chain.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_goldens.py tests/test_gpu_model.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
python profiles/run_flow_stage.py 16 5 > gpurun_out/r02_f_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chainup -c 1 -o gpurun_out/r02_f_chainup python profiles/run_flow_stage.py 16 5 > gpurun_out/r02_f_ncu.log 2>&1
tail -2 gpurun_out/r02_f_ncu.log
