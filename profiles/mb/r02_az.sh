#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --batch 32 --grid 28 --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_az.json 2> gpurun_out/r02_az.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_az.json").read().strip().splitlines()[-1])
print("g28 ms_per_step", round(d["ms_per_step"], 4))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"]):
    if "conv" in k: print("   %-32s %7.3f ms x%.0f" % (k, v["ms_per_step"], v["launches_per_step"]))
PY
