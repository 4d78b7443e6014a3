#!/bin/bash
# DDP pre-training step at N GPUs, FastSyncBatchNorm vs torch.nn.SyncBatchNorm (both with the graphed momentum branch)
N=${1:-2}
mkdir -p gpurun_out
for mode in "" "--no-fast-syncbn"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 \
    main_pretrain.py --synthetic --batch-size 128 --n-frames 6 --steps 10 --warmup 5 --print-freq 100 $mode 2> gpurun_out/r02_o_n${N}.err | tail -1 | cut -c1-330 | tee -a gpurun_out/r02_o_pretrain_n${N}.jsonl
done
tail -3 gpurun_out/r02_o_n${N}.err
