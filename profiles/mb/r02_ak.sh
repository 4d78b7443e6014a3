#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --sparse --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_ak_sparse$i.json 2> gpurun_out/r02_ak_sparse$i.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_ak_sparse$i.json").read().strip().splitlines()[-1])
print("sparse run $i ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4))
PY
done
