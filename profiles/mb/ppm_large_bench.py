"""PPM fwd+bwd at the large grids (14x14, 28x28): tcgen05 3xTF32 route vs the CUDA-core fp32 route
(PIXPRO_B200_NO_TC=1) vs the reference ops in eager PyTorch (fp32, TF32 off)."""
import os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops, _cabi
torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
def torch_ppm(x, v):
    N, C, H, W = x.shape
    vh = F.normalize(v, dim=1).view(N, C, -1); xh = F.normalize(x, dim=1).view(N, C, -1)
    att = torch.clamp(torch.bmm(xh.transpose(1, 2), xh), min=0) ** 2
    return F.normalize(torch.bmm(vh, att.transpose(1, 2)).view(N, C, H, W), dim=1)
for G in (14, 28):
    x = torch.randn(B, 256, G, G, device="cuda", requires_grad=True); v = torch.randn(B, 256, G, G, device="cuda", requires_grad=True)
    g = torch.randn(B, 256, G, G, device="cuda")
    def ours():
        x.grad = v.grad = None
        ops.ppm(x, v, 2.0, 0.0, True).backward(g)
    def ref():
        x.grad = v.grad = None
        torch_ppm(x, v).backward(g)
    flops = 2 * 10 * (G * G) ** 2 * 256 * B  # (4+6) P^2 C MACs per sample-view
    ms = t(ours); msr = t(ref)
    print(f"G={G} B={B}: ours {ms:.3f} ms ({flops/ms/1e9:.1f} TFLOP/s useful)   eager torch fp32 {msr:.3f} ms   [NO_TC={os.environ.get('PIXPRO_B200_NO_TC','0')}]")
