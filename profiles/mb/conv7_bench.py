"""Value transform at the 7x7 grid (1x1 conv 256 -> 256 on [128, 256, 7, 7]) through this package's kernels: device time of every
launch (the library's event brackets).  History: written for the A/B of the padded TMA route (PIXPRO_B200_CONVPAD, removed after
profiles/r02_v_conv7.txt / r02_w_convpad_step.txt); it now times the thread-staged kernels the 7x7 grid runs on."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import _cabi, ops  # noqa: E402

B, C, G = 128, 256, 7
ops.set_serial(True) if hasattr(ops, "set_serial") else None
x = torch.randn(B, C, G, G, device="cuda", requires_grad=True)
w = (torch.randn(C, C, 1, 1, device="cuda") / 16).requires_grad_(True)
b = torch.zeros(C, device="cuda", requires_grad=True)
g = torch.randn(B, C, G, G, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step():
    x.grad = w.grad = b.grad = None
    ops.conv1x1(x, w, b).backward(g)


for _ in range(5):
    step()
torch.cuda.synchronize()
_cabi.profile_enable(True)
for _ in range(20):
    flush.zero_()
    step()
torch.cuda.synchronize()
rep = _cabi.profile_report()
_cabi.profile_enable(False)
print("route:", "thread-staged" if os.environ.get("PIXPRO_B200_CONVPAD") == "0" else "padded TMA")
for k, (l, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:<32} {ms / l * 1000:7.1f} us/launch x {l / 20:.0f}")
