#!/usr/bin/env python
"""main_pretrain.py — the reference's pre-training entry (main_pretrain.py:40-272) around this repo's drop-in modules.

Same structure as the reference: `build_model` (model from `contrast.models`, LARS(SGD) from `contrast.lars`,
DistributedDataParallel with broadcast_buffers=False, :40-85), `train_step` = one iteration of `train()` (:207-270:
host->device of the loader tuple, `util.apply_optical_flow`, mask-ratio logging, `model(...)`, zero_grad / backward /
optimizer.step / scheduler.step).  One process per GPU under torchrun, NCCL gradient all-reduce.

What is different, deliberately:
  * `--synthetic`: the loader is replaced by seeded synthetic batches in the loader's tuple layout
    (contrast/data/dataset.py:503: [img, img2, coord, coord2, index, [target, flow_fwd, flow_bwd], [size, num_img]]).
    The reference's data pipeline, option parser, logger, scheduler and checkpointing are out of scope of this repo;
    without --synthetic this script needs the reference tree on sys.path behind this package and
    uses ITS `contrast.data.get_loader` / `contrast.option.parse_option` / `contrast.lr_scheduler.get_scheduler`.
  * the reference's `size, cur_n_frames = info; util.calc_frame_ratio(...)` (:229-231) references a function that does
    not exist in its own util.py; the intended bookkeeping (:232-241) is kept, the dead call is not reproduced.
  * per-iteration `.item()` host reads (loss, pos_num, mask ratio; :243-251,273-290) happen only every
    `print_freq` iterations: they are logging, and each one drains the GPU.

  python main_pretrain.py --synthetic --batch-size 128 --n-frames 6 --steps 20              # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 main_pretrain.py --synthetic ...
"""
import argparse
import json
import math
import os
import sys
import time
import types

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pixpro-with-opticalflow_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

from contrast import models, resnet, util  # noqa: E402
from contrast.lars import LARS, add_weight_decay  # noqa: E402
from pixpro_b200 import _cabi, synth  # noqa: E402


def synthetic_options(**kw):
    """The reference's option namespace (contrast/option.py) restricted to what this entry reads, at the published
    settings (tools/pretrain_bdd100k_job_base.sh:249-282)."""
    a = types.SimpleNamespace(
        arch="resnet50", model="PixPro", batch_size=128, n_frames=6, image_size=224, epochs=2000, start_epoch=1, warmup_epoch=5,
        optimizer="lars", base_learning_rate=1.0, weight_decay=1e-5, momentum=0.9, amp="bf16", amp_opt_level="O0",
        pixpro_p=2.0, pixpro_momentum=0.99, pixpro_pos_ratio=0.7, pixpro_clamp_value=0.0, pixpro_transform_layer=1,
        pixpro_ins_loss_weight=0.0, feature_dim=256, head_type="early_return", output_dir="/tmp", num_instances=70000,
        use_flow=True, use_flow_file=True, use_flow_frames=False, flow_up=True, flow_cat_norm=False, alpha1=0.01, alpha2=0.5,
        flow_sparse=False, graph_momentum_branch=True, fast_sync_bn=True, channels_last=True, debug=False, print_freq=100, local_rank=0,
        frame_hw=(720, 1280), flow_hw=(90, 160), steps_per_epoch=500)
    a.__dict__.update(kw)
    a.use_flow = a.use_flow and a.n_frames > 1
    return a


def build_model(args, device):
    """main_pretrain.py:40-85 (RAFT on the fly is out of scope: flows come precomputed, use_flow_file)."""
    encoder = resnet.__dict__[args.arch]
    model = models.__dict__[args.model](encoder, args).to(device)
    model.graph_momentum_branch = bool(getattr(args, "graph_momentum_branch", False))
    lr = args.batch_size * dist.get_world_size() / 256 * args.base_learning_rate
    if args.optimizer == "sgd":
        optimizer = torch.optim.SGD(model.parameters(), lr=lr, momentum=args.momentum, weight_decay=args.weight_decay)
    elif args.optimizer == "lars":
        optimizer = LARS(torch.optim.SGD(add_weight_decay(model, args.weight_decay), lr=lr, momentum=args.momentum))
    else:
        raise NotImplementedError
    if args.use_flow and not args.use_flow_file:
        raise NotImplementedError("this synthetic entry feeds precomputed links (--use_flow_file); for on-the-fly estimation build "
                                  "contrast.flow.RAFT, load the reference's checkpoint and pass it to util.apply_optical_flow "
                                  "as flow_model (main_pretrain.py:44-58 of the reference)")
    ddp = DistributedDataParallel(model, device_ids=[device.index], broadcast_buffers=False)
    return ddp, optimizer


class CosineWithWarmup:
    """lr schedule of the published runs (contrast/lr_scheduler.py: linear warm-up over warmup_epoch, then cosine),
    stepped every iteration — host arithmetic only."""

    def __init__(self, optimizer, n_iter_per_epoch, args):
        self.opt, self.n, self.args = optimizer, n_iter_per_epoch, args
        self.base = [g["lr"] for g in optimizer.param_groups]
        self.t = 0

    def step(self):
        self.t += 1
        warm = self.args.warmup_epoch * self.n
        total = self.args.epochs * self.n
        if self.t < warm:
            f = self.t / max(warm, 1)
        else:
            f = 0.5 * (1.0 + math.cos(math.pi * (self.t - warm) / max(total - warm, 1)))
        for g, b in zip(self.opt.param_groups, self.base):
            g["lr"] = b * f


class SyntheticLoader:
    """Seeded batches in the loader's tuple layout, kept in PINNED host memory like a DataLoader(pin_memory=True)
    would deliver them (`resident=True`: already on the device, for the kernel-side number)."""

    def __init__(self, args, device, rank, n_batches=2, resident=True):
        B, n = args.batch_size, max(args.n_frames - 1, 1)
        H, W = args.frame_hw
        self.batches = []
        for i in range(n_batches):
            g = torch.Generator().manual_seed(100 + 17 * rank + i)
            im1 = torch.randn(B, 3, args.image_size, args.image_size, generator=g)
            im2 = torch.randn(B, 3, args.image_size, args.image_size, generator=g)
            if args.channels_last:
                im1, im2 = im1.contiguous(memory_format=torch.channels_last), im2.contiguous(memory_format=torch.channels_last)
            c1 = synth.crop_coords(B, W, H, seed=1 + 10 * rank + i)
            c2 = synth.crop_coords(B, W, H, seed=2 + 10 * rank + i)
            batch = [im1, im2, c1, c2, torch.arange(B)]
            if args.use_flow:
                lo_f, lo_b = synth.flow_fields(B, n, h=args.flow_hw[0], w=args.flow_hw[1], seed=3 + rank + i)
                batch += [[torch.zeros(B, dtype=torch.long), lo_f, lo_b], [torch.tensor([[H, W]] * B), torch.tensor([[args.n_frames]] * B)]]
            else:
                batch += [torch.zeros(B, dtype=torch.long), [torch.tensor([[H, W]] * B), torch.tensor([[1]] * B)]]
            place = (lambda t: t.to(device)) if resident else (lambda t: t.pin_memory())
            self.batches.append([[place(x) for x in item] if isinstance(item, list) else place(item) for item in batch])
        self.i = 0

    def __len__(self):
        return 1 << 30

    def next(self):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        return b


def train_step(data, model, optimizer, scheduler, args, log=None):
    """One iteration of the reference's train() (main_pretrain.py:207-270)."""
    data = [[x.cuda(non_blocking=True) for x in item] if isinstance(item, (tuple, list)) else item.cuda(non_blocking=True)
            for item in data]                                                           # :209-216
    is_mask_flow = args.use_flow and args.alpha1 is not None and args.alpha2 is not None
    if args.use_flow:
        flow_fwd, flow_bwd = util.apply_optical_flow(data, None, args)                  # :226
        mask_fwd, mask_bwd = flow_fwd[2], flow_bwd[2]
        data[2] = [data[2], flow_fwd]                                                   # :240-241
        data[3] = [data[3], flow_bwd]
    want_log = log is not None
    if is_mask_flow and want_log:                                                       # :243-251 (logging only)
        r_fwd, r_bwd = util.calc_mask_ratio(mask_fwd).mean(), util.calc_mask_ratio(mask_bwd).mean()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.amp == "bf16"):
        loss, pos_num_list = model(data[0], data[1], data[2], data[3])                  # :259
    optimizer.zero_grad()                                                               # :262
    loss.backward()                                                                     # :267 (DDP all-reduce overlapped)
    optimizer.step()                                                                    # :268
    scheduler.step()                                                                    # :269
    if want_log:
        (pn1, pm1), (pn2, pm2) = pos_num_list
        log.update(loss=float(loss.item()), pos_num=float(pn1.sum().item() + pn2.sum().item()),
                   lr=optimizer.param_groups[0]["lr"])
        if is_mask_flow:
            log["mask_ratio"] = float((r_fwd + r_bwd).item() / 2.0)
    return loss


class SyntheticTrainer:
    """What bench.py --pretrain times: build_model + SyntheticLoader + train_step."""

    def __init__(self, args, device):
        torch.manual_seed(0)  # identical initial weights on every rank (DDP also broadcasts them)
        self.args, self.dev = args, device
        rank = dist.get_rank()
        self.model, self.optimizer = build_model(args, device)
        self.model.train()
        self.scheduler = CosineWithWarmup(self.optimizer, args.steps_per_epoch, args)
        self.loader = SyntheticLoader(args, device, rank)
        self.flow_mode = "none" if not args.use_flow else ("sparse correspondence" if args.flow_sparse else "dense (pp_flow_stage)")
        self._n0 = _cabi.launch_count()
        self._steps = 0

    def step(self, log=None):
        self._steps += 1
        return train_step(self.loader.next(), self.model, self.optimizer, self.scheduler, self.args, log)

    def launches(self):
        """Kernels of THIS repo's library launched so far (the backbone's cuDNN / ATen launches are not counted)."""
        return _cabi.launch_count() - self._n0

    def breakdown(self, reps=5):
        """Device time of the repo's own parts of one step, run alone: flow stage, optimizer step."""
        a, data = self.args, self.loader.next()
        out = {}

        def timed(fn):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        if a.use_flow:
            out["flow_stage"] = timed(lambda: util.apply_optical_flow(data, None, a))
        out["lars_sgd_step"] = timed(self.optimizer.step)
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--n-frames", type=int, default=6)
    ap.add_argument("--amp", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--flow-sparse", action="store_true")
    ap.add_argument("--no-graph-momentum-branch", action="store_true")
    ap.add_argument("--no-fast-syncbn", action="store_true", help="torch.nn.SyncBatchNorm instead of pixpro_b200.syncbn.FastSyncBatchNorm")
    ap.add_argument("--print-freq", type=int, default=10)
    a = ap.parse_args()
    if not a.synthetic:
        raise SystemExit("main_pretrain.py: only --synthetic is self-contained.  For real data put the reference tree on sys.path "
                         "behind this package (INTEGRATION.md §1) and run the reference's own main_pretrain.py: its imports of "
                         "contrast.models / contrast.util / contrast.flow / contrast.lars then resolve to this package.")
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29571")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    args = synthetic_options(batch_size=a.batch_size, n_frames=a.n_frames, amp=a.amp, flow_sparse=a.flow_sparse,
                             graph_momentum_branch=not a.no_graph_momentum_branch, fast_sync_bn=not a.no_fast_syncbn,
                             print_freq=a.print_freq, local_rank=local)
    tr = SyntheticTrainer(args, dev)
    for _ in range(a.warmup):
        tr.step()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    log = {}
    for i in range(a.steps):
        tr.step(log if (i + 1) % a.print_freq == 0 else None)
        if rank == 0 and (i + 1) % a.print_freq == 0:
            print(f"Train: [{i + 1}/{a.steps}] lr {log['lr']:.3f} loss {log['loss']:.3f} pos_num {log['pos_num']:.0f} "
                  f"mask ratio {log.get('mask_ratio', float('nan')):07.3%} T {(time.perf_counter() - t0) / (i + 1):.3f}s/it", file=sys.stderr)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "PixPro+OF pretrain frames/sec", "value": a.batch_size * world * a.n_frames / ms * 1e3, "unit": "frames/s",
                          "n_gpus": world, "ms_per_step": ms, "per_gpu_batch": a.batch_size, "n_frames": a.n_frames, "amp": a.amp,
                          "flow_stage": tr.flow_mode, "steps": a.steps, "warmup": a.warmup, "data": "synthetic"}))
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)  # a CUDA graph holding NCCL kernels must not outlive an orderly communicator teardown (it hangs)


if __name__ == "__main__":
    main()
