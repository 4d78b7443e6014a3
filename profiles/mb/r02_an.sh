#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_an_bench.json 2> gpurun_out/r02_an_bench.err
timeout 300 python bench.py --sparse --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_an_sparse.json 2> gpurun_out/r02_an_sparse.err
python - <<'PY'
import json
for f in ("r02_an_bench", "r02_an_sparse"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4))
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"]):
        if "small" in k: print("   %-32s %7.3f ms x%.0f" % (k, v["ms_per_step"], v["launches_per_step"]))
PY
