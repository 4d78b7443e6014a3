"""EMA of the key branch and the LARS+SGD step on the ResNet-50 + projector parameter set (the model of
main_pretrain.py): fused multi-tensor kernels vs the reference's per-parameter torch loops on the same GPU."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from contrast import resnet  # noqa: E402
from contrast.lars import LARS, add_weight_decay  # noqa: E402
from pixpro_b200 import _cabi, optim  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
online = resnet.resnet50(head_type="early_return").to(dev)
key = resnet.resnet50(head_type="early_return").to(dev)
n = sum(p.numel() for p in online.parameters())
pairs = [(q.data, k.data) for q, k in zip(online.parameters(), key.parameters())]
print(f"{len(pairs)} tensors, {n / 1e6:.1f} M parameters")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) / reps * 1e3


m = 0.9931


def ema_ref():
    for q, k in pairs:
        k.copy_(k * m + q * (1. - m))


gpu, wall = timed(lambda: optim.ema_update(pairs, m, cache_key="bench"))
print(f"EMA fused : {gpu * 1e3:8.1f} us device, {wall * 1e3:8.1f} us wall, {12 * n / gpu / 1e6:7.0f} GB/s (12 B/element)")
gpu, wall = timed(ema_ref)
print(f"EMA torch loop (reference's way): {gpu * 1e3:8.1f} us device, {wall * 1e3:8.1f} us wall")

for p in online.parameters():
    p.grad = torch.randn_like(p) * 1e-3
opt = LARS(torch.optim.SGD(add_weight_decay(online, 1e-5), lr=0.1, momentum=0.9))
gpu, wall = timed(opt.step)
print(f"LARS+SGD fused: {gpu * 1e3:8.1f} us device, {wall * 1e3:8.1f} us wall, {28 * n / gpu / 1e6:7.0f} GB/s (28 B/element)")


def lars_ref(groups, state, trust=0.001, eps=1e-8):  # the reference's per-parameter sequence, with its host syncs
    with torch.no_grad():
        for g in groups:
            for p in g["params"]:
                grad = p.grad.add(p, alpha=g["weight_decay"]) if g["weight_decay"] > 0 else p.grad
                if not g["ignore"]:
                    pn, gn = p.norm(), grad.norm()
                    a = 1.0
                    if pn > 0 and gn > 0:
                        a = trust * pn / (gn + eps)
                    grad = grad.mul(a)
                buf = state.setdefault(p, torch.zeros_like(p))
                buf.mul_(0.9).add_(grad)
                p.add_(buf, alpha=-g["lr"])


st = {}
gpu, wall = timed(lambda: lars_ref(opt.param_groups, st), reps=5)
print(f"LARS+SGD per-parameter torch ops (reference's way): {gpu * 1e3:8.1f} us device, {wall * 1e3:8.1f} us wall")
