#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_z_tests.log; cat gpurun_out/r02_z_tests.log
timeout 600 python bench.py --batch 32 --grid 28 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_z_g28.json 2> gpurun_out/r02_z_g28.err
timeout 600 python bench.py --batch 64 --grid 14 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_z_g14.json 2> gpurun_out/r02_z_g14.err
python - <<'PY'
import json
for f in ("g28", "g14"):
    try:
        d = json.loads(open("gpurun_out/r02_z_%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4))
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"]):
        print("   %-32s %7.3f ms x%.0f hbm %s tensor %s" % (k, v["ms_per_step"], v["launches_per_step"], v.get("hbm_frac"), v.get("tensor_frac")))
PY
