"""Minimal driver for ncu captures of the tcgen05 3xTF32 GEMM (PPM 28x28 similarity shape)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops
A = torch.randn(64, 784, 256, device="cuda"); B = torch.randn(64, 784, 256, device="cuda")
for _ in range(3):
    C = ops.tc_gemm_nt(A, B)
torch.cuda.synchronize()
print("ok", C[0, 0, 0].item())
