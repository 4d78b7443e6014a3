#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_j_tests.log; cat gpurun_out/r02_j_tests.log
grep -q " passed" gpurun_out/r02_j_tests.log && ! grep -q "failed" gpurun_out/r02_j_tests.log || exit 1
timeout 600 python bench.py --batch 32 --grid 28 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_j_g28.json 2> gpurun_out/r02_j_g28.err; tail -2 gpurun_out/r02_j_g28.err
timeout 600 python bench.py --batch 64 --grid 14 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_j_g14.json 2> gpurun_out/r02_j_g14.err; tail -2 gpurun_out/r02_j_g14.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_j_g28.json", "gpurun_out/r02_j_g14.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms_per_step", d["ms_per_step"], "e2e", d.get("e2e", {}).get("ms_per_step"))
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"]):
        print("   %-40s %7.3f ms x%.0f tensor %s" % (k, v["ms_per_step"], v["launches_per_step"], v.get("tensor_frac")))
PY
