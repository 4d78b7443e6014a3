"""Layout probe for the MN-major operand path of the TMA-fed GEMM: B = identity rows (K-major), A(m,k) = m + 1000 k stored
MN-major; C[m][n<K] then shows which A element the tensor core read at (m, k = n)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops
torch.set_printoptions(linewidth=250, sci_mode=False)
M, N, K = 128, 256, 16
m = torch.arange(M).view(M, 1).float()
k = torch.arange(K).view(1, K).float()
A = (m + 1000 * k)[None].cuda()                 # [1, M, K]
B = torch.zeros(1, N, K).cuda()
for i in range(K):
    B[0, i, i] = 1.0
for a_mn, b_mn in [(False, False), (True, False), (False, True)]:
    Ain = A.transpose(1, 2).contiguous() if a_mn else A
    Bin = B.transpose(1, 2).contiguous() if b_mn else B
    C = ops.tc_gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn)[0].cpu()
    want = A[0].cpu()
    got = C[:, :K]
    print(f"a_mn={a_mn} b_mn={b_mn}: exact match {bool(torch.equal(got, want))}; nonzero outside first K columns: {int((C[:, K:] != 0).sum())}")
    if not torch.equal(got, want):
        gm, gk = (got % 1000).long(), (got // 1000).long()
        print(" rows (m') read at m = 0..7, k = 0..15:\n", gm[:8])
        print(" k' read at m = 0..7, k = 0..15:\n", gk[:8])
        print(" rows read at m = 32..35:\n", gm[32:36], "\n k':\n", gk[32:36])
        print(" full C rows 0..1 first 32 columns:\n", C[:2, :32])
