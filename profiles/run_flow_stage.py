"""Minimal driver for ncu captures of the flow-stage kernels: B samples, n links, 3 repetitions."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
f, b = synth.flow_fields(B, n, seed=1)
f, b = f.cuda(), b.cuda()
for _ in range(3):
    out = ops.flow_stage(f, b)
torch.cuda.synchronize()
print("ok", out[2].float().mean().item())
