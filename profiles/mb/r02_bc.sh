#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_final_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"], 4), round(d["e2e"]["value"]), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), "sparse", round(d["sparse_correspondence"]["ms_per_step"], 4))
for k, v in d["configs"].items(): print(k, round(v["ms_per_step"], 4))
print("pretrain", d["pretrain_ddp"]["ms_per_step"])
PY
