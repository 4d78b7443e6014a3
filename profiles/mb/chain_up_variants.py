"""Fused x8-up-sampling chain kernel (pp_chainup.cuh) variants, PIXPRO_B200_CHAINUP=1..3: device time of the
"chain_up" launch at B samples, n links, output checksum, slow-path pixel-links per launch."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import _cabi, ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
f, b = synth.flow_fields(B, n, seed=1)
f, b = f.cuda(), b.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    out = ops.flow_stage(f, b)
torch.cuda.synchronize()
_cabi.chain_slow_count(reset=True)
_cabi.profile_enable(True)
for _ in range(10):
    flush.zero_()
    out = ops.flow_stage(f, b)
torch.cuda.synchronize()
rep = _cabi.profile_report()
_cabi.profile_enable(False)
slow = _cabi.chain_slow_count()
h = hashlib.sha256(out[0].cpu().numpy().tobytes() + out[1].cpu().numpy().tobytes()).hexdigest()[:16]
print(f"variant={os.environ.get('PIXPRO_B200_CHAINUP', 'default')} B={B} n={n}", {k: round(ms / l * 1000, 1) for k, (l, ms) in rep.items()},
      f"flow sha {h}  slow pixel-links/launch {slow / 10:.0f} of {B * 2 * 720 * 1280 * n}")
