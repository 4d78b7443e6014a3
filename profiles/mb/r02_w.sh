#!/bin/bash
# historical: PIXPRO_B200_CONVPAD selected the padded TMA route of the 7x7 value transform, removed after these measurements (profiles/r02_w_convpad_step.txt)
mkdir -p gpurun_out
for pad in 1 0; do
  PIXPRO_B200_CONVPAD=$pad timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_w_pad$pad.json 2> gpurun_out/r02_w_pad$pad.err
  PIXPRO_B200_CONVPAD=$pad timeout 300 python bench.py --sparse --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_w_sparse_pad$pad.json 2>> gpurun_out/r02_w_pad$pad.err
done
python - <<'PY'
import json
for f in ("r02_w_pad1", "r02_w_pad0", "r02_w_sparse_pad1", "r02_w_sparse_pad0"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4), "sparse", d.get("sparse_correspondence", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r02_w_pad1.err
