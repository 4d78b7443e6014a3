#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_ah_tests.log; cat gpurun_out/r02_ah_tests.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_ah_bench.json 2> gpurun_out/r02_ah_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_ah_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4), "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("traffic"))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"])[:5]:
    print("   %-32s %7.3f ms x%.0f hbm %s" % (k, v["ms_per_step"], v["launches_per_step"], v.get("hbm_frac")))
PY
