"""GPU: the tcgen05 (3xTF32) batched GEMM against torch float64 — numerics test of a floating
kernel: tolerance 1e-5 relative to the result's scale (the path's fp32 bar), written here."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("tma", [False, True], ids=["staged", "tma"])
@pytest.mark.parametrize("batch,M,N,K", [(1, 128, 128, 32), (2, 128, 128, 256), (3, 196, 196, 256), (2, 256, 784, 784),
                                         (1, 50, 70, 36), (2, 300, 130, 100), (2, 784, 784, 256), (1, 130, 530, 72)])
def test_tc_gemm_nt_matches_fp64(batch, M, N, K, tma):
    """tma=True: the TMA-fed warp-specialised kernel (csrc/pp_tc2.cuh; 128 x 256 tiles, ragged M / N / K tails are TMA
    zero fill); tma=False: the thread-staged kernel (csrc/pp_tc.cuh)."""
    from pixpro_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    A = torch.randn(batch, M, K, generator=g).to(DEV)
    B = torch.randn(batch, N, K, generator=g).to(DEV)
    C = ops.tc_gemm_nt(A, B, tma=tma)
    ref = torch.bmm(A.double(), B.double().transpose(1, 2))
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err
    # and it is genuinely better than single-pass TF32 would be (~1e-3)
    assert err < 5e-6


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("batch,M,N,K", [(2, 128, 256, 64), (3, 196, 196, 256), (2, 784, 784, 256), (2, 256, 784, 784), (1, 132, 520, 72)])
def test_tc_gemm_operand_layouts(batch, M, N, K, a_mn, b_mn):
    """The TMA-fed kernel with operands in place in either layout (K-major [rows, K] or MN-major [K, rows]: TMA boxes of
    32 rows x 16 k-lines and MN-major UMMA descriptors), fp32 operands split by the kernel's converter warps."""
    from pixpro_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + K)
    A = torch.randn(batch, M, K, generator=g).to(DEV)
    B = torch.randn(batch, N, K, generator=g).to(DEV)
    Ain = A.transpose(1, 2).contiguous() if a_mn else A
    Bin = B.transpose(1, 2).contiguous() if b_mn else B
    C = ops.tc_gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn)
    ref = torch.bmm(A.double(), B.double().transpose(1, 2))
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 5e-6, err


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, True), (True, False), (False, True)])
@pytest.mark.parametrize("batch,M,N,K,kb", [(5, 256, 49, 256, 1), (9, 256, 256, 49, 4), (128, 256, 256, 49, 4), (3, 100, 25, 60, 2),
                                            (2, 130, 300, 21, 1)])
def test_tc_gemm_padded_and_batch_sums(batch, M, N, K, kb, a_mn, b_mn):
    """pp_tc_gemm_ex: operands padded to a pitch of 4 floats along their contiguous dimension (the pad holds garbage that must
    never be read: it is NaN here) and products summed over groups of kb batch entries (the last group may be short)."""
    from pixpro_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + K + kb)
    A = torch.randn(batch, M, K, generator=g).to(DEV)
    B = torch.randn(batch, N, K, generator=g).to(DEV)

    def padded(t, mn):  # [batch, rows, K] -> its stored form with the last dimension padded to a multiple of 4 (+4: a real pad)
        t = t.transpose(1, 2).contiguous() if mn else t
        pitch = ((t.shape[2] + 3) // 4) * 4 + 4
        out = torch.full((t.shape[0], t.shape[1], pitch), float("nan"), device=DEV)
        out[:, :, :t.shape[2]] = t
        return out

    C = ops.tc_gemm_padded(padded(A, a_mn), padded(B, b_mn), M, N, K, a_mn=a_mn, b_mn=b_mn, kb=kb)
    ref = torch.bmm(A.double(), B.double().transpose(1, 2))
    groups = (batch + kb - 1) // kb
    ref = torch.stack([ref[i * kb:(i + 1) * kb].sum(0) for i in range(groups)])
    assert C.shape == ref.shape
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 5e-6, err


@pytest.mark.parametrize("tma", [False, True], ids=["staged", "tma"])
def test_tc_gemm_exact_on_small_integers(tma):
    """Integer-valued operands below 2^10 are exact in TF32 and sums below 2^24 are exact in
    fp32: the tensor-core result must equal the integer matmul bit for bit."""
    from pixpro_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randint(-8, 9, (2, 200, 64), generator=g).float().to(DEV)
    B = torch.randint(-8, 9, (2, 136, 64), generator=g).float().to(DEV)
    C = ops.tc_gemm_nt(A, B, tma=tma)
    ref = torch.bmm(A.double(), B.double().transpose(1, 2)).float()
    assert torch.equal(C, ref)


@pytest.mark.parametrize("B,Cin,Cout,G", [(128, 256, 256, 7), (5, 256, 256, 14), (3, 64, 96, 5), (7, 32, 64, 7), (2, 8, 12, 3), (6, 256, 256, 13)])
def test_conv1x1_matches_torch(B, Cin, Cout, G):
    """value transform (1x1 conv) forward/backward on the tensor cores vs cuDNN in true fp32."""
    import torch.nn.functional as F
    from pixpro_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + G)
    x = torch.randn(B, Cin, G, G, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(Cout, Cin, 1, 1, generator=g) / 16).to(DEV).requires_grad_(True)
    b = torch.randn(Cout, generator=g).to(DEV).requires_grad_(True)
    dy = torch.randn(B, Cout, G, G, generator=g).to(DEV)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(x.double(), w.double(), b.double())
        gx, gw, gb = torch.autograd.grad(ref, (x, w, b), dy.double())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    y = ops.conv1x1(x, w, b)
    y.backward(dy)
    for got, want in ((y, ref), (x.grad, gx), (w.grad, gw), (b.grad, gb)):
        err = (got.double() - want).abs().max().item() / want.abs().max().item()
        assert err < 1e-5, err
