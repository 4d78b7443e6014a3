"""Microbenchmark: tcgen05 3xTF32 batched GEMM vs torch.bmm fp32 (TF32 off) at the PPM shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for name, batch, M, N, K in [("S 14x14", 64, 196, 196, 256), ("S 28x28", 64, 784, 784, 256), ("Y 28x28", 64, 256, 784, 784), ("big", 8, 2048, 2048, 2048)]:
    A = torch.randn(batch, M, K, device="cuda"); B = torch.randn(batch, N, K, device="cuda")
    from pixpro_b200 import _cabi
    ws = torch.empty(_cabi.lib().pp_tc_gemm_nt_workspace(batch, M, N, K), dtype=torch.uint8, device="cuda")
    ms_old = t(lambda: ops.tc_gemm_nt(A, B, tma=False))
    ms_tc = t(lambda: ops.tc_gemm_nt(A, B, tma=True, workspace=ws)); ms_th = t(lambda: torch.bmm(A, B.transpose(1, 2)))
    fl = 2.0 * batch * M * N * K
    _cabi.profile_enable(True)
    for _ in range(5): ops.tc_gemm_nt(A, B, tma=True, workspace=ws)
    torch.cuda.synchronize(); rep = _cabi.profile_report(); _cabi.profile_enable(False)
    kern = {k: v[1] / v[0] for k, v in rep.items()}
    print(f"{name:8s} thread-staged kernel {ms_old:.3f} ms = {fl/ms_old/1e9:.1f} TFLOP/s | TMA-fed kernel alone {kern.get('tc2_gemm_nt', float('nan')):.3f} ms = "
          f"{fl/kern.get('tc2_gemm_nt', float('nan'))/1e9:.1f} TFLOP/s, hi/lo split launches {kern.get('tc split', float('nan')):.3f} ms each")
    C = ops.tc_gemm_nt(A, B, tma=True, workspace=ws); ref = torch.bmm(A.double(), B.double().transpose(1, 2))
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    e2 = ((torch.bmm(A, B.transpose(1, 2)).double() - ref).abs().max() / ref.abs().max()).item()
    print(f"{name:8s} [{batch}x{M}x{N}x{K}] tcgen05 3xTF32: {ms_tc:.3f} ms = {fl/ms_tc/1e9:.1f} TFLOP/s (err {err:.1e}) | torch fp32 bmm: {ms_th:.3f} ms = {fl/ms_th/1e9:.1f} TFLOP/s (err {e2:.1e})")
