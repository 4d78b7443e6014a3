#!/bin/bash
# historical: PIXPRO_B200_PPMREG=0 still selects the two-pass kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py tests/test_gpu_reference_goldens.py -x -q -k "ppm or featprop or model or pixpro" 2>&1 | tail -6 > gpurun_out/r02_y_tests.log; cat gpurun_out/r02_y_tests.log
for reg in 1 0; do
  PIXPRO_B200_PPMREG=$reg timeout 600 python bench.py --batch 32 --grid 28 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_y_g28_reg$reg.json 2> gpurun_out/r02_y_g28_reg$reg.err
  PIXPRO_B200_PPMREG=$reg timeout 600 python bench.py --batch 64 --grid 14 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_y_g14_reg$reg.json 2> gpurun_out/r02_y_g14_reg$reg.err
done
python - <<'PY'
import json
for f in ("g28_reg1", "g28_reg0", "g14_reg1", "g14_reg0"):
    try:
        d = json.loads(open("gpurun_out/r02_y_%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4))
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"]):
        if "tcgen05" not in k: print("   %-32s %7.3f ms x%.0f hbm %s" % (k, v["ms_per_step"], v["launches_per_step"], v.get("hbm_frac")))
PY
