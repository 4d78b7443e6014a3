import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import _cabi, ops, synth
f, b = synth.flow_fields(2, 1, seed=1)
f, b = f.cuda(), b.cuda()
try:
    out = ops.flow_stage(f, b)
    torch.cuda.synchronize()
    print("ran ok", out[2].float().mean().item())
except Exception as e:
    print("ERR", str(e)[:200])
print("redo", _cabi.fb_redo_count())
