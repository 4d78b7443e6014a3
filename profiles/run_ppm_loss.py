"""Minimal driver for ncu captures of the PPM + loss kernels: B samples, GxG grid, fwd+bwd, 3 reps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
G = int(sys.argv[2]) if len(sys.argv) > 2 else 7
f, b = synth.flow_fields(B, 1, seed=1)
ff, fb, mf, mb = ops.flow_stage(f.cuda(), b.cuda())
feat1, feat2, k1, k2 = [t.cuda() for t in synth.features(B, 256, G, seed=2)]
c1, c2 = synth.crop_coords(B, seed=3).cuda(), synth.crop_coords(B, seed=4).cuda()
for _ in range(3):
    x = feat1.clone().requires_grad_(True)
    pred = ops.ppm(x, x, 2.0, 0.0, True)
    loss, pn, pm = ops.regression_loss(pred, k2, c1, c2, 0.7, flow=ff, size=(720, 1280), mask=mf)
    loss.backward()
torch.cuda.synchronize()
print("ok", loss.item())
