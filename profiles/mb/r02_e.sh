#!/bin/bash
# round 2, call E: fused up-sampling chain kernel v3 (taps from T8) — variants, parity, ncu capture
mkdir -p gpurun_out
for v in 1 2 3 4; do PIXPRO_B200_CHAINUP=$v python profiles/mb/chain_up_variants.py 64 5; done > gpurun_out/r02_e_chain.txt 2>&1
PIXPRO_B200_CHAINUP=1 python profiles/mb/chain_up_variants.py 64 2 >> gpurun_out/r02_e_chain.txt 2>&1
cat gpurun_out/r02_e_chain.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py tests/test_gpu_reference_goldens.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
python profiles/run_flow_stage.py 16 5 > gpurun_out/r02_e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chainup -c 1 -o gpurun_out/r02_e_chainup python profiles/run_flow_stage.py 16 5 > gpurun_out/r02_e_ncu.log 2>&1
tail -3 gpurun_out/r02_e_ncu.log
