#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_al_bench_${N}gpu.json 2> gpurun_out/r02_al_bench_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_al_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), round(d["e2e"]["value"]), "pretrain ms", round(d["pretrain_ddp"]["ms_per_step"], 2), round(d["pretrain_ddp"]["value"]))
PY
nvidia-smi topo -m | head -12
tail -2 gpurun_out/r02_al_bench_${N}gpu.err
