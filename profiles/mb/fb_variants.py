"""FB-mask kernel variants (PIXPRO_B200_FBTILE=0: gather kernel, 1..4: TMA-staged tile kernel with
different box / stage counts): device time of the "fb" launch, mask checksum, fix-up pixel count."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
os.environ.setdefault("PIXPRO_B200_FBUP", "0")  # this script studies fbbox_kernel itself: the two-kernel route
from pixpro_b200 import _cabi, ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
f, b = synth.flow_fields(B, n, seed=1)
f, b = f.cuda(), b.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    out = ops.flow_stage(f, b)
torch.cuda.synchronize()
_cabi.fb_redo_count(reset=True)
_cabi.profile_enable(True)
for _ in range(10):
    flush.zero_()
    out = ops.flow_stage(f, b)
torch.cuda.synchronize()
rep = _cabi.profile_report()
_cabi.profile_enable(False)
redo = _cabi.fb_redo_count()
h = hashlib.sha256(out[2].cpu().numpy().tobytes() + out[3].cpu().numpy().tobytes()).hexdigest()[:16]
l, ms = rep["fb"] if "fb" in rep else (sum(rep[k][0] for k in ("fb_up_w", "fb1", "fb_up") if k in rep) // 2,
                                        sum(rep[k][1] for k in ("fb_up_w", "fb1", "fb_up") if k in rep))  # fused route: both mask launches
print(f"variant={os.environ.get('PIXPRO_B200_FBTILE', 'default')} fb {ms / l * 1000:.1f} us/launch  mask sha {h}  "
      f"valid {out[2].float().mean().item():.4f}  redo pixels/launch {redo / 10:.0f}")
