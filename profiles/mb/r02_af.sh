#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -4
for up in 2 1 0; do
  PIXPRO_B200_FBUP=$up timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_af_up$up.json 2> gpurun_out/r02_af_up$up.err
done
PIXPRO_B200_FBUP=1 timeout 300 python bench.py --sparse --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_af_sparse.json 2> gpurun_out/r02_af_sparse.err
python - <<'PY'
import json
for f in ("r02_af_up2", "r02_af_up1", "r02_af_up0", "r02_af_sparse"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4))
PY
