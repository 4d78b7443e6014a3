"""Whole pre-training step under DDP on N GPUs of one box (BASELINE.json metric (i): PixPro+OF pretrain frames/sec
@1/2/4/8 B200).  One process per GPU (torchrun), NCCL gradient all-reduce over NVLink, SyncBatchNorm in the encoders /
projectors as in the reference (contrast/models/PixPro.py:289-292, main_pretrain.py:78), this package's drop-in
modules for everything on the pixel path:

  flow stage : contrast.util.apply_optical_flow (dense, or --sparse: lazy stand-ins + pp_sparse_corr inside the loss)
  model      : contrast.models.PixPro forward + backward (ResNet-50 + projectors on PyTorch/cuDNN as north_star
               prescribes; PPM, value transform, flow-guided loss, EMA on the sm_100a kernels)
  optimizer  : contrast.lars.LARS(SGD) (multi-tensor kernels)

  python profiles/mb/pretrain_ddp.py [--batch 64] [--n-frames 2] [--amp bf16|fp32] [--channels-last] [--sparse]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/mb/pretrain_ddp.py ...

Synthetic data resident on the GPU (no loader).  Step time = CUDA events around `steps` steps, max over ranks.
Prints one JSON line on rank 0.  Weak scaling: the per-GPU batch is fixed."""
import argparse
import json
import os
import sys
import types

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from contrast import resnet, util  # noqa: E402
from contrast.lars import LARS, add_weight_decay  # noqa: E402
from contrast.models import PixPro  # noqa: E402
from pixpro_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--n-frames", type=int, default=2)
ap.add_argument("--amp", default="bf16", choices=["bf16", "fp32"])
ap.add_argument("--channels-last", action="store_true")
ap.add_argument("--sparse", action="store_true")
ap.add_argument("--graph-key-branch", action="store_true", help="momentum branch replayed from one CUDA graph")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=4)
a = ap.parse_args()

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29571")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
torch.manual_seed(0)  # identical initial weights on every rank (DDP also broadcasts them)
B, n = a.batch, a.n_frames - 1
margs = types.SimpleNamespace(pixpro_p=2.0, pixpro_momentum=0.99, pixpro_pos_ratio=0.7, pixpro_clamp_value=0.0,
                              pixpro_transform_layer=1, pixpro_ins_loss_weight=0.0, output_dir="/tmp", num_instances=100000,
                              batch_size=B, epochs=100, start_epoch=1, feature_dim=256, head_type="early_return")
model = PixPro(resnet.resnet50, margs).to(dev)
model.graph_momentum_branch = a.graph_key_branch
opt = LARS(torch.optim.SGD(add_weight_decay(model, 1e-5), lr=B * world / 256 * 1.0, momentum=0.9))
ddp = DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False)  # main_pretrain.py:78
g = torch.Generator(device="cpu").manual_seed(100 + rank)
im1, im2 = (torch.randn(B, 3, 224, 224, generator=g).to(dev) for _ in range(2))
if a.channels_last:
    im1, im2 = im1.contiguous(memory_format=torch.channels_last), im2.contiguous(memory_format=torch.channels_last)
c1, c2 = synth.crop_coords(B, seed=1 + 10 * rank).to(dev), synth.crop_coords(B, seed=2 + 10 * rank).to(dev)
use_flow = n >= 1
if use_flow:
    lo_f, lo_b = (t.to(dev) for t in synth.flow_fields(B, n, seed=3 + rank))
    fargs = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=True, flow_up=True,
                                  flow_cat_norm=False, debug=False, flow_sparse=a.sparse)
    data = [None] * 7
    data[5] = [None, lo_f, lo_b]
    data[6] = [torch.tensor([[720, 1280]] * B), torch.tensor([[a.n_frames]] * B)]
amp = a.amp == "bf16"


def step():
    coord1, coord2 = c1, c2
    if use_flow:
        f1, f2 = util.apply_optical_flow(data, None, fargs)  # main_pretrain.py:226
        coord1, coord2 = [c1, f1], [c2, f2]
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        loss, _ = ddp(im1, im2, coord1, coord2)               # :259
    opt.zero_grad()
    loss.backward()                                           # :267 (DDP all-reduce overlapped with the backward)
    opt.step()
    return loss


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
lv = loss.detach().float().clone()
dist.all_reduce(lv)
if rank == 0:
    print(json.dumps({"metric": "PixPro+OF pretrain step (flow stage + ResNet-50 x2 branches fwd/bwd + DDP all-reduce + LARS), frames/sec",
                      "value": B * world * a.n_frames / ms * 1e3, "unit": "frames/s", "n_gpus": world, "ms_per_step": ms,
                      "samples_per_s": B * world / ms * 1e3, "per_gpu_batch": B, "n_frames": a.n_frames, "amp": a.amp,
                      "channels_last": a.channels_last, "flow_stage": "sparse" if a.sparse else "dense", "graph_key_branch": a.graph_key_branch, "scaling": "weak",
                      "grad_allreduce_mb": sum(p.numel() for p in model.parameters() if p.requires_grad) * 4 / 1e6,
                      "mean_loss": float(lv.item()) / world, "steps": a.steps, "warmup": a.warmup, "data": "synthetic"}))
torch.cuda.synchronize()
dist.barrier()
if a.graph_key_branch:
    # a CUDA graph that holds NCCL kernels must not outlive an orderly communicator teardown (it hangs): leave at once
    sys.stdout.flush()
    os._exit(0)
dist.destroy_process_group()
