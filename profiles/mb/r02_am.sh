#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --batch 32 --grid 28 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_am_g28.json 2> gpurun_out/r02_am_g28.err
timeout 600 python bench.py --batch 64 --grid 14 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_am_g14.json 2> gpurun_out/r02_am_g14.err
python - <<'PY'
import json
for f in ("g28", "g14"):
    d = json.loads(open("gpurun_out/r02_am_%s.json" % f).read().strip().splitlines()[-1])
    print(f, "ms_per_step", round(d["ms_per_step"], 4))
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"]):
        if "loss" in k: print("   %-32s %7.3f ms x%.0f" % (k, v["ms_per_step"], v["launches_per_step"]))
PY
