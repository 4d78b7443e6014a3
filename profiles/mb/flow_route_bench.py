"""Flow stage (n = 1, masks) as one pp_flow_stage call at several batch sizes, CUDA-graph replay, L2 flushed per repetition:
device time per call for the route PIXPRO_B200_FBUP selects (0: upchain1 + fbbox; 1: upchain1(fwd) + fbbox_up_w + fbbox_up;
2: upchain1(fwd) + fbbox_up_w + fbbox on the forward direction)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops, synth  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = []
for B in (8, 16, 32, 64):
    f, b = synth.flow_fields(B, 1, seed=1)
    f, b = f.cuda(), b.cuda()
    out = (torch.empty((B, 2, 720, 1280), device="cuda"), torch.empty((B, 2, 720, 1280), device="cuda"),
           torch.empty((B, 720, 1280), device="cuda", dtype=torch.uint8), torch.empty((B, 720, 1280), device="cuda", dtype=torch.uint8))
    for _ in range(3):
        ops.flow_stage(f, b, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.flow_stage(f, b, out=out)
    ts = []
    for _ in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    res.append(f"B={B}: {ts[len(ts) // 2] * 1000:.1f} us ({ts[len(ts) // 2] * 1000 / B:.2f} us/sample)")
print("FBUP=" + os.environ.get("PIXPRO_B200_FBUP", "default") + "  " + "   ".join(res))
