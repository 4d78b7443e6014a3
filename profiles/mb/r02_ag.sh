#!/bin/bash
mkdir -p gpurun_out
for up in 0 1 2; do PIXPRO_B200_FBUP=$up timeout 200 python profiles/mb/flow_route_bench.py; done > gpurun_out/r02_ag_flow_routes.txt 2>&1
cat gpurun_out/r02_ag_flow_routes.txt
