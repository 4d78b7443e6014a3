"""value_transform (1x1 conv 256->256 on [2B,256,7,7]) fwd+bwd: cuDNN conv2d vs cuBLAS matmul, fp32 (TF32 off)."""
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1000
B, C, G = 128, 256, 7
x = torch.randn(B, C, G, G, device="cuda", requires_grad=True); w = (torch.randn(C, C, 1, 1, device="cuda") / 16).requires_grad_(True)
b = torch.zeros(C, device="cuda", requires_grad=True); g = torch.randn(B, C, G, G, device="cuda")
def conv():
    x.grad = w.grad = b.grad = None
    F.conv2d(x, w, b).backward(g)
def mm():
    x.grad = w.grad = b.grad = None
    (torch.matmul(w.view(C, C), x.flatten(2)) + b.view(1, C, 1)).view(B, C, G, G).backward(g)
def mm2():  # one big GEMM over (b,p): x -> [C, B*P]
    x.grad = w.grad = b.grad = None
    xt = x.flatten(2).permute(1, 0, 2).reshape(C, -1)
    y = torch.addmm(b.view(C, 1), w.view(C, C), xt).view(C, B, G * G).permute(1, 0, 2).reshape(B, C, G, G)
    y.backward(g)
print(f"conv2d fwd+bwd {t(conv):.1f} us | matmul(bmm) {t(mm):.1f} us | single GEMM with permutes {t(mm2):.1f} us")
y1 = F.conv2d(x, w, b); y2 = (torch.matmul(w.view(C, C), x.flatten(2)) + b.view(1, C, 1)).view(B, C, G, G)
print("max diff", (y1 - y2).abs().max().item())
