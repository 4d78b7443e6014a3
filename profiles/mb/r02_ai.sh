#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_ai_bench_2gpu.json 2> gpurun_out/r02_ai_bench_2gpu.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_ai_ref_2gpu.json 2> gpurun_out/r02_ai_ref_2gpu.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_ai_bench_2gpu.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "pretrain", d["pretrain_ddp"]["ms_per_step"], d["pretrain_ddp"]["value"])
r = json.loads(open("gpurun_out/r02_ai_ref_2gpu.json").read().strip().splitlines()[-1])
print("ref", r["impl"], r["value"], r["n_gpus"])
PY
tail -2 gpurun_out/r02_ai_bench_2gpu.err
