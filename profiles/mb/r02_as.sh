#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for up in 0 2; do PIXPRO_B200_FBUP=$up timeout 200 python profiles/mb/flow_route_bench.py; done
timeout 300 python bench.py --steps 60 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_as_bench.json 2> gpurun_out/r02_as_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_as_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"])[:4]:
    print("   %-32s %7.3f ms x%.0f hbm %s" % (k, v["ms_per_step"], v["launches_per_step"], v.get("hbm_frac")))
PY
