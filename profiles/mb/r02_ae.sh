#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py tests/test_gpu_reference_goldens.py -x -q -k "flow or fb or bench or stage" 2>&1 | tail -6
for up in 2 1; do
  PIXPRO_B200_FBUP=$up timeout 300 python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_ae_up$up.json 2> gpurun_out/r02_ae_up$up.err
done
python - <<'PY'
import json
for f in ("r02_ae_up2", "r02_ae_up1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4), "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
    for k, v in sorted(d.get("kernels", {}).items(), key=lambda x: -x[1]["ms_per_step"])[:4]:
        print("   %-32s %7.3f ms x%.0f hbm %s" % (k, v["ms_per_step"], v["launches_per_step"], v.get("hbm_frac")))
PY
