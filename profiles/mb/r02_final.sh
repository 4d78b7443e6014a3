#!/bin/bash
# round-2 final measurement pass (one B200): bench line, reference arm, ncu launch list of the bench command, --set full captures
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/r02_final_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_final_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbbox -c 2 -o gpurun_out/r02_final_fbbox \
    python profiles/run_flow_stage.py 64 1 > gpurun_out/r02_final_ncu_fb.log 2>&1; echo "ncu fb rc=$?"   # fbbox_up_kernel<WRITE> and fbbox_kernel (one direction)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:colnorm_scale|normbwd_reg" -c 4 -o gpurun_out/r02_final_ppmreg \
    python bench.py --grid 28 --batch 32 --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-graph > gpurun_out/r02_final_ncu_ppm.log 2>&1; echo "ncu ppm rc=$?"
ls -la gpurun_out/r02_final*
