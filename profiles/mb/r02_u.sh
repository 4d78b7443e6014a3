#!/bin/bash
mkdir -p gpurun_out
{
echo "# FBTILE default (4 CTAs/SM)"; timeout 300 python profiles/mb/flow_overlap.py 64 1 4,8,16,32
echo "# FBTILE=2 (3 CTAs/SM, 72 regs)"; PIXPRO_B200_FBTILE=2 timeout 300 python profiles/mb/flow_overlap.py 64 1 4,8,16,32
} > gpurun_out/r02_u_flow_overlap.txt 2>&1
cat gpurun_out/r02_u_flow_overlap.txt
