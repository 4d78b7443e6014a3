#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "full_size_batch_sample or threshold_margin or corr" -s 2>&1 | tail -15 > gpurun_out/r02_l_tests.log; cat gpurun_out/r02_l_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_l_bench.json 2> gpurun_out/r02_l_bench.err; tail -3 gpurun_out/r02_l_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_l_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roofline frac", d["roofline"]["frac"])
print("sparse", d.get("sparse_correspondence", {}).get("ms_per_step"), "decoupled", d.get("decoupled_dense"))
print("pretrain", {k: d.get("pretrain_ddp", {}).get(k) for k in ("value", "ms_per_step", "error")})
for k, v in d.get("configs", {}).items():
    print(k, v.get("ms_per_step"), v.get("error"))
PY
