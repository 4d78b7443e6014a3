"""TMA-fed GEMM operand modes at the PPM shapes: pre-split planes (PIXPRO_B200_TC2_PLANES=1 through pp_tc_gemm_nt_ws), fp32 operands
split in the kernel (K-major), and MN-major operands read in place."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops, _cabi


def kern_ms(fn, name, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    _cabi.profile_enable(True)
    for _ in range(n): fn()
    torch.cuda.synchronize(); rep = _cabi.profile_report(); _cabi.profile_enable(False)
    return {k: v[1] / v[0] for k, v in rep.items()}.get(name, float("nan"))


for name, batch, M, N, K in [("S 28x28", 64, 784, 784, 256), ("Y 28x28", 64, 256, 784, 784), ("conv 28x28", 64, 256, 784, 256)]:
    A = torch.randn(batch, M, K, device="cuda"); B = torch.randn(batch, N, K, device="cuda")
    At, Bt = A.transpose(1, 2).contiguous(), B.transpose(1, 2).contiguous()
    fl = 2.0 * batch * M * N * K
    res = {}
    if os.environ.get("PIXPRO_B200_TC2_PLANES") == "1":
        ws = torch.empty(_cabi.lib().pp_tc_gemm_nt_workspace(batch, M, N, K), dtype=torch.uint8, device="cuda")
        res["planes"] = kern_ms(lambda: ops.tc_gemm_nt(A, B, tma=True, workspace=ws), "tc2_gemm_nt")
    else:
        res["raw K/K"] = kern_ms(lambda: ops.tc_gemm(A, B), "tc2_gemm")
        res["raw MN/K"] = kern_ms(lambda: ops.tc_gemm(At, B, a_mn=True), "tc2_gemm")
        res["raw K/MN"] = kern_ms(lambda: ops.tc_gemm(A, Bt, b_mn=True), "tc2_gemm")
        res["raw MN/MN"] = kern_ms(lambda: ops.tc_gemm(At, Bt, a_mn=True, b_mn=True), "tc2_gemm")
    print(f"{name:10s} [{batch}x{M}x{N}x{K}] " + " | ".join(f"{k}: {v:.3f} ms = {fl / v / 1e9:.0f} TFLOP/s" for k, v in res.items()))
