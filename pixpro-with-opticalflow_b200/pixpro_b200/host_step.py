"""Host-buffer entry of the pixel-pretext hot path: pinned host tensors in, pinned host tensors out.

This is the end-to-end call a data-loader-side user makes (and what bench.py's `e2e` times):
the loader's flow links / crop descriptors and the feature maps live in pinned host memory;
`HostPixelStep` stages them to the device, runs flow stage -> PPM -> paired loss -> backward, and
returns loss / positive counts / feature gradients in pinned host memory.

Copies and kernels are overlapped with two CUDA streams instead of being serialised: the flow
links go first on the compute stream (the flow kernels need only them), while features, keys and
crop descriptors are copied on a side stream underneath the flow kernels; the compute stream waits
on an event before the PPM.  Nothing here computes: all arithmetic is in libpixpro_b200.so.
"""
import torch
import torch.nn.functional as F

from . import ops


class HostPixelStep:
    def __init__(self, device, batch, channels=256, grid=7, size=(720, 1280), gamma=2.0, clamp=0.0, pos_ratio=0.7,
                 alpha1=0.01, alpha2=0.5, flow_up=True):
        self.dev = torch.device(device)
        self.size, self.gamma, self.clamp, self.pos_ratio = size, gamma, clamp, pos_ratio
        self.alpha1, self.alpha2, self.flow_up = alpha1, alpha2, flow_up
        self.side = torch.cuda.Stream(device=self.dev)
        self.ready = torch.cuda.Event()
        self.out = {"loss": torch.empty((), dtype=torch.float32).pin_memory(),
                    "pos_num": torch.empty((2, batch), dtype=torch.float32).pin_memory(),
                    "d_feat": torch.empty((2, batch, channels, grid, grid), dtype=torch.float32).pin_memory()}

    def h2d_bytes(self, host):
        return sum(t.numel() * t.element_size() for k, t in host.items() if k not in ("w", "bias"))

    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.out.values())

    def __call__(self, host, w, bias):
        """host: dict of pinned tensors lo_f, lo_b (optional), feat1, feat2, k1, k2, c1, c2.
        w, bias: value_transform parameters (device-resident, they are model state)."""
        dev = self.dev
        main = torch.cuda.current_stream(dev)
        use_flow = "lo_f" in host
        if use_flow:  # first in the queue: the flow kernels depend on nothing else
            lo_f = host["lo_f"].to(dev, non_blocking=True)
            lo_b = host["lo_b"].to(dev, non_blocking=True)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            t = {k: host[k].to(dev, non_blocking=True) for k in ("feat1", "feat2", "k1", "k2", "c1", "c2")}
            self.ready.record(self.side)
        ff = fb = mf = mb = None
        if use_flow:
            ff, fb, mf, mb = ops.flow_stage(lo_f, lo_b, flow_up=self.flow_up, alpha_1=self.alpha1, alpha_2=self.alpha2)
        main.wait_event(self.ready)
        for v in t.values():
            v.record_stream(main)
        f12 = torch.cat([t["feat1"], t["feat2"]], dim=0).requires_grad_(True)
        wg = w.detach().requires_grad_(True)
        bg = bias.detach().requires_grad_(True)
        pred1, pred2 = ops.ppm(f12, F.conv2d(f12, wg, bg), self.gamma, self.clamp, final_norm=True).chunk(2, dim=0)
        l12, pn, _ = ops.regression_loss_pair(pred1, t["k2"], t["c1"], t["c2"], pred2, t["k1"], t["c2"], t["c1"],
                                              self.pos_ratio, flow1=ff, flow2=fb, size=self.size, mask1=mf, mask2=mb)
        loss = l12[0] + l12[1]
        loss.backward()
        B = t["feat1"].shape[0]
        self.out["loss"].copy_(loss.detach(), non_blocking=True)
        self.out["pos_num"].copy_(pn, non_blocking=True)
        self.out["d_feat"].copy_(f12.grad.view(2, B, *f12.shape[1:]), non_blocking=True)
        main.synchronize()  # the caller reads the loss every step
        return self.out, (wg.grad, bg.grad)
