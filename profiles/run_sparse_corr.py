"""Minimal driver for ncu captures of the sparse correspondence kernel: B samples, n links, G x G grid."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
G = int(sys.argv[3]) if len(sys.argv) > 3 else 7
f, b = synth.flow_fields(B, n, seed=1)
f, b = f.cuda(), b.cuda()
c1, c2 = synth.crop_coords(B, seed=2).cuda(), synth.crop_coords(B, seed=3).cuda()
for _ in range(3):
    wf, wb = ops.sparse_corr(f, b, c1, c2, G, (720, 1280))
torch.cuda.synchronize()
print("ok", wf[2].mean().item(), wb[2].mean().item())
