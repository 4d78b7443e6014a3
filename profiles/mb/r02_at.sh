#!/bin/bash
mkdir -p gpurun_out
for flag in "" "--no-overlap"; do
timeout 300 python bench.py --steps 60 --warmup 5 --no-extra --no-cpu-baseline $flag > gpurun_out/r02_at.json 2> gpurun_out/r02_at.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_at.json").read().strip().splitlines()[-1])
print("flag '$flag' ms_per_step", round(d["ms_per_step"], 4))
PY
done
