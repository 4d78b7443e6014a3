#!/bin/bash
# the driver's multi-GPU invocation of bench.py (both arms), N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_r_bench_n${N}.json 2> gpurun_out/r02_r_bench_n${N}.err
echo "exit $?"; tail -3 gpurun_out/r02_r_bench_n${N}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_r_bench_n${N}.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
print("pretrain", {k: d.get("pretrain_ddp", {}).get(k) for k in ("value", "ms_per_step", "n_gpus", "error")})
print("configs", {k: v.get("ms_per_step", v.get("error")) for k, v in d.get("configs", {}).items()})
print("cpu_baseline", d.get("cpu_baseline", {}).get("value"))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29592 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
