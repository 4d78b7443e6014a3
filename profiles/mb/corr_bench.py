"""Microbenchmark: RAFT correlation volume + pyramid + 12 windowed lookups (csrc/pp_corr.cu through contrast.flow.corr.CorrBlock)
against the torch-op sequence the reference's CorrBlock issues on the same GPU (contrast/flow/corr.py:12-60, restated here as
test infrastructure: matmul, avg_pool2d, per-level meshgrid + grid_sample + cat + permute)."""
import os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from contrast.flow.corr import CorrBlock
torch.backends.cuda.matmul.allow_tf32 = False


class TorchCorrBlock:
    def __init__(self, f1, f2, num_levels=4, radius=4):
        B, D, h, w = f1.shape
        corr = torch.matmul(f1.view(B, D, h * w).transpose(1, 2), f2.view(B, D, h * w)).view(B, h, w, 1, h, w) / torch.sqrt(torch.tensor(D).float())
        corr = corr.reshape(B * h * w, 1, h, w)
        self.pyr, self.L, self.r = [corr], num_levels, radius
        for _ in range(num_levels - 1):
            corr = F.avg_pool2d(corr, 2, stride=2)
            self.pyr.append(corr)

    def __call__(self, coords):
        r = self.r
        coords = coords.permute(0, 2, 3, 1)
        B, h, w, _ = coords.shape
        out = []
        for i in range(self.L):
            corr = self.pyr[i]
            d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
            delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), axis=-1)
            c = coords.reshape(B * h * w, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
            H, W = corr.shape[-2:]
            grid = torch.cat([2 * c[..., :1] / (W - 1) - 1, 2 * c[..., 1:] / (H - 1) - 1], dim=-1)
            out.append(F.grid_sample(corr, grid, align_corners=True).view(B, h, w, -1))
        return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n


for name, B, D, h, w, r in [("RAFT-small 368x496", 8, 128, 46, 62, 3), ("RAFT-basic 368x496", 8, 256, 46, 62, 4), ("RAFT-small 720x1280", 2, 128, 90, 160, 3)]:
    g = torch.Generator().manual_seed(1)
    f1, f2 = torch.randn(B, D, h, w, generator=g).cuda(), torch.randn(B, D, h, w, generator=g).cuda()
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    coords = (torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1) + 3 * torch.randn(B, 2, h, w, generator=g)).cuda()
    def run(cls):
        blk = cls(f1, f2, num_levels=4, radius=r)
        for _ in range(12):
            o = blk(coords)
        return o
    ms_mine, ms_torch = t(lambda: run(CorrBlock)), t(lambda: run(TorchCorrBlock))
    a, b = run(CorrBlock), run(TorchCorrBlock)
    err = ((a - b).abs().max() / b.abs().max()).item()
    vol_gb = B * (h * w) ** 2 * 4 / 1e9
    print(f"{name:22s} B={B}: volume {vol_gb:.2f} GB; build + 12 lookups: this repo {ms_mine:.3f} ms | torch ops {ms_torch:.3f} ms ({ms_torch / ms_mine:.1f}x); max rel diff {err:.1e}")
