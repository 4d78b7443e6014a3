#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_aa_tests.log; cat gpurun_out/r02_aa_tests.log
grep -q " passed" gpurun_out/r02_aa_tests.log && ! grep -q "failed\|error" gpurun_out/r02_aa_tests.log || exit 1
timeout 600 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_aa_bench.json 2> gpurun_out/r02_aa_bench.err
timeout 600 python bench.py --sparse --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_aa_sparse.json 2>> gpurun_out/r02_aa_bench.err
python - <<'PY'
import json
for f in ("r02_aa_bench", "r02_aa_sparse"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4), "sparse", d.get("sparse_correspondence", {}).get("ms_per_step"), "launches/step", d.get("gpu_launches_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r02_aa_bench.err
