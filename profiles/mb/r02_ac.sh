#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_ac_tests.log; cat gpurun_out/r02_ac_tests.log
for up in 1 0; do
  PIXPRO_B200_FBUP=$up timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_ac_up$up.json 2> gpurun_out/r02_ac_up$up.err
done
python - <<'PY'
import json
for f in ("r02_ac_up1", "r02_ac_up0"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4), "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"])[:6]:
        print("   %-32s %7.3f ms x%.0f hbm %s" % (k, v["ms_per_step"], v["launches_per_step"], v.get("hbm_frac")))
PY
tail -3 gpurun_out/r02_ac_up1.err
python -c "
import sys; sys.path.insert(0,'pixpro-with-opticalflow_b200')
from pixpro_b200 import _cabi
print('redo', _cabi.fb_redo_count())"
