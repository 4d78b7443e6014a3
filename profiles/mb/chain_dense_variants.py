"""Dense-link chain (concat_flow, n links of [B,2,720,1280]) : gather kernel vs TMA-staged kernel
(PIXPRO_B200_CHAINBOX=0/1), device time per launch and output checksum."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import _cabi, ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
f, _ = synth.flow_fields(B, n, seed=3)
up = ops.upflow8(f.cuda().reshape(-1, 2, 90, 160)).reshape(B, n, 2, 720, 1280).permute(1, 0, 2, 3, 4).contiguous()
for _ in range(2):
    out = ops.concat_flow(up)
torch.cuda.synchronize()
_cabi.profile_enable(True)
for _ in range(5):
    out = ops.concat_flow(up)
torch.cuda.synchronize()
rep = _cabi.profile_report()
_cabi.profile_enable(False)
h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
print(os.environ.get("PIXPRO_B200_CHAINBOX", "default"), {k: round(ms / l * 1000, 1) for k, (l, ms) in rep.items()}, "sha", h)
