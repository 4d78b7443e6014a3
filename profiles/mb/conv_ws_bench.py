"""A/B of the tcgen05 GEMM kernels (PIXPRO_B200_TCWS=0: CTA-synchronous, 1: warp-specialised ring) on the value-transform
1x1 conv at the bench shape (2B=128 samples, 256->256, 7x7): forward, dgrad, wgrad."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import _cabi, ops
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1000
B, C, G = 128, 256, 7
x = torch.randn(B, C, G, G, device="cuda"); w = torch.randn(C, C, device="cuda") / 16; bias = torch.zeros(C, device="cuda")
dy = torch.randn(B, C, G, G, device="cuda"); y = torch.empty_like(x); dx = torch.empty_like(x); dw = torch.empty_like(w)
L = _cabi.lib(); st = torch.cuda.current_stream().cuda_stream
ws = torch.empty(L.pp_conv1x1_bwd_workspace(B, C, C, G * G), dtype=torch.uint8, device="cuda")
p = lambda a: a.data_ptr()
fwd = lambda: L.pp_conv1x1_fwd(p(x), p(w), p(bias), B, C, C, G * G, p(y), st)
dgr = lambda: L.pp_conv1x1_bwd(p(x), p(w), p(dy), B, C, C, G * G, p(dx), None, None, p(ws), st)
wgr = lambda: L.pp_conv1x1_bwd(p(x), p(w), p(dy), B, C, C, G * G, None, p(dw), None, p(ws), st)
print(f"TCWS={os.environ.get('PIXPRO_B200_TCWS', '1')}: conv1x1 fwd {t(fwd):.1f} us | dgrad {t(dgr):.1f} us | wgrad (+reduce) {t(wgr):.1f} us")
ref = torch.einsum("oc,bchw->bohw", w.double(), x.double())
print("fwd rel err", ((y.double() - ref).abs().max() / ref.abs().max()).item())
