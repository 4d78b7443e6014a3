"""GPU: FastSyncBatchNorm (csrc/pp_syncbn.cu: three launches and one all-reduce per direction) against torch's batch norm —
the arithmetic torch.nn.SyncBatchNorm performs on one rank (contrast/models/PixPro.py:289-292 converts every BatchNorm of
the model to it).  fp32: 1e-5 relative to the result's scale; bf16 activations: one bf16 ulp of the result's scale.
Two simulated ranks are covered on the CPU side (tests/test_host_logic.py, gloo, the kernels restated in torch)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pair(C, affine=True, momentum=0.1):
    from pixpro_b200.syncbn import FastSyncBatchNorm
    ref = nn.BatchNorm2d(C, momentum=momentum, affine=affine).to(DEV)
    fast = FastSyncBatchNorm(C, momentum=momentum, affine=affine).to(DEV)
    if affine:
        with torch.no_grad():
            ref.weight.uniform_(0.5, 1.5)
            ref.bias.uniform_(-0.5, 0.5)
    fast.load_state_dict(ref.state_dict())
    return ref, fast


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("shape", [(8, 64, 56, 56), (4, 256, 14, 14), (3, 33, 7, 5), (16, 2048, 7, 7), (5, 48, 1, 1)])
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_backward_match_torch_batch_norm(shape, channels_last, dtype):
    N, C, H, W = shape
    ref, fast = _pair(C)
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(shape, generator=g) * 1.7 + 0.6).to(DEV).to(dtype)
    dy = torch.randn(shape, generator=g).to(DEV).to(dtype)
    if channels_last:
        x, dy = x.contiguous(memory_format=torch.channels_last), dy.contiguous(memory_format=torch.channels_last)
    xr = x.detach().clone().requires_grad_(True)
    xf = x.detach().clone().requires_grad_(True)
    yr = ref(xr)
    yf = fast(xf)
    assert yf.dtype == x.dtype and yf.shape == x.shape
    yr.backward(dy)
    yf.backward(dy)
    tol = 1e-5 if dtype == torch.float32 else 1.0 / 128   # bf16: 8 bits of mantissa
    assert _rel(yf, yr) < tol
    assert _rel(xf.grad, xr.grad) < (5e-5 if dtype == torch.float32 else 1.0 / 64)
    stat_tol = 1e-5 if dtype == torch.float32 else 1e-3
    assert _rel(fast.weight.grad, ref.weight.grad) < (1e-4 if dtype == torch.float32 else 2e-2)
    assert _rel(fast.bias.grad, ref.bias.grad) < (1e-4 if dtype == torch.float32 else 2e-2)
    assert _rel(fast.running_mean, ref.running_mean) < stat_tol
    assert _rel(fast.running_var, ref.running_var) < stat_tol
    assert int(fast.num_batches_tracked) == 1


def test_large_mean_channels_keep_their_variance():
    """mean >> deviation: the statistics are formed about a pivot, not as E[x^2] - E[x]^2 of raw fp32 sums."""
    ref, fast = _pair(32)
    x = (torch.randn(16, 32, 28, 28) * 0.01 + 50.0).to(DEV)
    yr, yf = ref(x), fast(x)
    assert _rel(fast.running_var, ref.running_var) < 5e-3
    assert ((yf - yr).abs().max() / yr.abs().max()).item() < 5e-3


def test_eval_mode_and_state_dict_are_the_parents():
    ref, fast = _pair(16)
    assert list(fast.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(4, 16, 8, 8, device=DEV)
    fast(x); ref(x)
    fast.eval(); ref.eval()
    assert torch.allclose(fast(x), ref(x), atol=1e-5, rtol=1e-5)


def test_conversion_replaces_every_batch_norm_in_place():
    from pixpro_b200.syncbn import FastSyncBatchNorm, convert_fast_sync_batchnorm
    net = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.ReLU(), nn.Sequential(nn.Conv2d(8, 4, 1), nn.SyncBatchNorm(4))).to(DEV)
    keys = list(net.state_dict().keys())
    w = net[1].weight
    out = convert_fast_sync_batchnorm(net)
    assert out is net and isinstance(net[1], FastSyncBatchNorm) and isinstance(net[3][1], FastSyncBatchNorm)
    assert net[1].weight is w and list(net.state_dict().keys()) == keys
    y = net(torch.randn(2, 3, 10, 10, device=DEV))
    y.square().mean().backward()
    assert net[1].weight.grad is not None and net[0].weight.grad is not None


def test_2d_input_and_no_affine():
    ref = nn.BatchNorm1d(24, affine=False).to(DEV)
    from pixpro_b200.syncbn import FastSyncBatchNorm
    fast = FastSyncBatchNorm(24, affine=False).to(DEV)
    x = torch.randn(64, 24, device=DEV, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    yr, yf = ref(x), fast(x2)
    dy = torch.randn_like(yr)
    yr.backward(dy); yf.backward(dy)
    assert _rel(yf, yr) < 1e-5 and _rel(x2.grad, x.grad) < 5e-5
