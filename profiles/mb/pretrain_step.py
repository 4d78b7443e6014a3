"""Whole pre-training step on one B200 (BASELINE.json configs[1] shape: ResNet-50, batch 64, 224x224 crops, 7x7 grid,
n_frames=2 flows 90x160 -> 720x1280), this package's drop-in modules vs the reference's own torch-op sequence
run on the same GPU (the restatements of tests/test_gpu_model.py; the reference tree itself is not on the GPU box).

  flow stage  : contrast.util.apply_optical_flow            vs  upflow8 + concat_flow + 2x FB consistency in torch ops
  model step  : contrast.models.PixPro forward + backward   vs  the same backbone with featprop / regression_loss in torch ops
  optimizer   : contrast.lars.LARS(SGD) + fused EMA         vs  per-parameter LARS / EMA loops
"""
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_model as T  # noqa: E402  (torch restatements of the reference's op sequence)
from contrast import resnet, util  # noqa: E402
from contrast.lars import LARS, add_weight_decay  # noqa: E402
from contrast.models import PixPro  # noqa: E402
from pixpro_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29561")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
amp = (sys.argv[2] if len(sys.argv) > 2 else "bf16") == "bf16"
torch.manual_seed(0)
model = PixPro(resnet.resnet50, T.pixpro_args(batch_size=B)).to(dev)
opt = LARS(torch.optim.SGD(add_weight_decay(model, 1e-5), lr=0.1, momentum=0.9))
im1, im2 = torch.randn(B, 3, 224, 224, device=dev), torch.randn(B, 3, 224, 224, device=dev)
c1, c2 = synth.crop_coords(B, seed=1).to(dev), synth.crop_coords(B, seed=2).to(dev)
lo_f, lo_b = (t.to(dev) for t in synth.flow_fields(B, 1, seed=3))
args = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=True, flow_up=True,
                             flow_cat_norm=False, debug=False)
data = [None] * 7
data[5] = [None, lo_f, lo_b]
data[6] = [torch.tensor([[720, 1280]] * B), torch.tensor([[2]] * B)]


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def flow_ours():
    with torch.no_grad():
        return util.apply_optical_flow(data, None, args)


def flow_torch():
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=False):
        ff, fb, mf, mb = T.torch_flow_stage(lo_f, lo_b)
    return [ff, (720, 1280), mf], [fb, (720, 1280), mb]


f1, f2 = flow_ours()


def step_ours():
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        loss, _ = model(im1, im2, [c1, f1], [c2, f2])
    loss.backward()


def step_torch():
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        loss, _, _ = T.torch_forward(model, im1, im2, c1, c2, f1, f2)
    loss.backward()


def ema_loop():
    with torch.no_grad():
        m = 0.99
        for q, k in zip(list(model.encoder.parameters()) + list(model.projector.parameters()),
                        list(model.encoder_k.parameters()) + list(model.projector_k.parameters())):
            k.copy_(k * m + q * (1. - m))


t_flow_o, t_flow_t = timed(flow_ours), timed(flow_torch, reps=3, warm=1)
t_step_o, t_step_t = timed(step_ours), timed(step_torch, reps=5, warm=2)
step_ours()
t_opt_o = timed(opt.step)
t_ema_t = timed(ema_loop, reps=5)
tot_o = t_flow_o + t_step_o + t_opt_o
print(f"B={B} amp={'bf16' if amp else 'fp32'} (the EMA update is inside the model step of this package)")
print(f"flow stage      : ours {t_flow_o:8.3f} ms   torch ops on the same GPU {t_flow_t:8.3f} ms   ({t_flow_t / t_flow_o:5.1f}x)")
print(f"model fwd+bwd   : ours {t_step_o:8.3f} ms   torch-op pixel path       {t_step_t:8.3f} ms   ({t_step_t / t_step_o:5.2f}x)")
print(f"LARS+SGD step   : ours {t_opt_o:8.3f} ms   (per-parameter EMA loop alone: {t_ema_t:.3f} ms)")
print(f"whole step      : ours {tot_o:8.3f} ms = {B * 2 / tot_o * 1e3:9.0f} frames/s on one B200 (n_frames=2)")
dist.destroy_process_group()
