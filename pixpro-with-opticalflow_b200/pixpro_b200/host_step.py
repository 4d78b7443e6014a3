"""Host-buffer entry of the pixel-pretext hot path: pinned host tensors in, pinned host tensors out.

This is the end-to-end call a data-loader-side user makes (and what bench.py's `e2e` times):
the loader's flow links / crop descriptors and the feature maps live in pinned host memory;
`HostPixelStep` stages them to the device, runs flow stage -> PPM -> paired loss -> backward, and
returns loss / positive counts / feature gradients in pinned host memory.

Copies and kernels are overlapped with two CUDA streams instead of being serialised: all host->
device copies are queued on a side stream, flow links first and in batch chunks of growing size
(B/8, 3B/8, B/2 by default: measured 0.933 ms/step against 0.959 ms with four equal chunks,
profiles/r02_bb_e2e_chunks.txt), then features, keys and crop descriptors.  The compute stream
starts the flow kernels of chunk i as soon as its event fires, so only the first chunk's copy (1/8
of the links) is exposed and the later, larger chunks run in the kernels' efficient regime;
everything else streams underneath the flow kernels.

With `use_graph=True` (default) the whole step — copies on both streams, every kernel of the path,
the autograd backward, the device->host copies — is captured ONCE into a CUDA graph and replayed:
per-step host work drops from ~50 Python-level launches to one cudaGraphLaunch, which is what the
end-to-end number is bound by once the kernels are fast.  The pinned input / output tensors are
the graph's fixed endpoints: the caller refills the same input tensors every step.
Nothing here computes: all arithmetic is in libpixpro_b200.so.
"""
import torch
import torch.nn.functional as F

from . import ops


class HostPixelStep:
    def __init__(self, device, batch, channels=256, grid=7, size=(720, 1280), gamma=2.0, clamp=0.0, pos_ratio=0.7,
                 alpha1=0.01, alpha2=0.5, flow_up=True, flow_chunks="auto", use_graph=True, sparse=False):
        self.dev = torch.device(device)
        self.size, self.gamma, self.clamp, self.pos_ratio = size, gamma, clamp, pos_ratio
        self.alpha1, self.alpha2, self.flow_up = alpha1, alpha2, flow_up
        # sparse=True: the flow stage is evaluated only at the loss's grid centres (ops.sparse_corr), the dense
        # composites / masks are never built; loss, counts and gradients are bit-identical to sparse=False
        self.sparse = sparse
        self.side = torch.cuda.Stream(device=self.dev)
        self.aux = torch.cuda.Stream(device=self.dev, priority=-1)  # few latency-bound blocks: first free SM slots
        self.ready = torch.cuda.Event()
        self.ready_feat = torch.cuda.Event()
        self.use_graph = use_graph
        self.graph = None
        self.calls = 0
        self.key = None
        self.param_grads = None
        # flow_chunks: an int = that many equal chunks; a list = the chunk sizes (must sum to `batch`); "auto" = progressive
        # sizes B/8, 3B/8, B/2: a small first chunk starts the flow kernels after 1/8 of the link copy, the later ones are
        # large enough for the kernels' efficient regime (>= 32 samples take the three-launch route, ops / pp_flow_stage)
        if sparse or batch < 16:
            sizes = [batch]
        elif flow_chunks == "auto":
            sizes = [batch // 8, 3 * batch // 8]
            sizes.append(batch - sum(sizes))
        elif isinstance(flow_chunks, (list, tuple)):
            sizes = [int(x) for x in flow_chunks]
            if sum(sizes) != batch or min(sizes) < 1:
                raise ValueError("flow_chunks: chunk sizes must be positive and sum to the batch")
        else:
            n = max(1, min(int(flow_chunks), batch))
            step = (batch + n - 1) // n
            sizes = [min(step, batch - b0) for b0 in range(0, batch, step)]
        self.chunk_sizes = sizes
        self.flow_chunks = len(sizes)
        self.chunk_ready = [torch.cuda.Event() for _ in range(self.flow_chunks)]
        self.out = {"loss": torch.empty((), dtype=torch.float32).pin_memory(),
                    "pos_num": torch.empty((2, batch), dtype=torch.float32).pin_memory(),
                    "d_feat": torch.empty((2, batch, channels, grid, grid), dtype=torch.float32).pin_memory()}

    def h2d_bytes(self, host):
        return sum(t.numel() * t.element_size() for k, t in host.items() if k not in ("w", "bias"))

    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.out.values())

    def __call__(self, host, w, bias):
        """host: dict of pinned tensors lo_f, lo_b (optional), feat1, feat2, k1, k2, c1, c2.
        w, bias: value_transform parameters (device-resident, they are model state).
        Returns (pinned outputs dict, (d_w, d_bias))."""
        if not self.use_graph:
            out = self._step(host, w, bias)
            torch.cuda.current_stream(self.dev).synchronize()  # the caller reads the loss every step
            return out
        key = tuple(sorted((k, v.data_ptr()) for k, v in host.items())) + (w.data_ptr(), bias.data_ptr())
        if self.graph is not None and key != self.key:
            raise ValueError("HostPixelStep(use_graph=True): pass the same pinned input tensors every step "
                             "(refill them in place); the captured graph copies from fixed addresses")
        self.calls += 1
        if self.graph is None:
            if self.calls <= 2:  # eager warm-up: lazy one-time setup (smem opt-ins, divisor certification, cuDNN plans)
                out = self._step(host, w, bias)
                torch.cuda.current_stream(self.dev).synchronize()
                return out
            self.key = key
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                _, self.param_grads = self._step(host, w, bias)
        self.graph.replay()
        torch.cuda.current_stream(self.dev).synchronize()
        return self.out, self.param_grads

    def _step(self, host, w, bias):
        dev = self.dev
        main = torch.cuda.current_stream(dev)
        use_flow = "lo_f" in host
        capturing = torch.cuda.is_current_stream_capturing()  # graph-private memory needs no record_stream
        self.side.wait_stream(main)
        chunks = []
        feat_keys, rest_keys = ("feat1", "feat2"), ("k1", "k2", "c1", "c2")
        # Copy queue (one PCIe direction, ~0.5 ms for the 27.6 MB of the bench batch — as long as the flow kernels): link
        # chunks first (the flow kernels are the critical path and wait for nothing else), then the two feature maps (all the
        # PPM forward needs: it starts before the keys have arrived), keys and crop descriptors last (the loss needs them).
        # Measured alternative: features right after the FIRST link chunk — the flow kernels then stall for the later chunks
        # and the step gets 1.5 % slower (profiles/r02_af_e2e_order.txt).
        with torch.cuda.stream(self.side):
            t = {}
            if use_flow and self.sparse:
                # sparse correspondence: the flow work is ~10 us, so the copy queue is ordered for the PPM alone —
                # features first, then keys / descriptors, links last
                t.update({k: host[k].to(dev, non_blocking=True) for k in feat_keys})
                self.ready_feat.record(self.side)
                t.update({k: host[k].to(dev, non_blocking=True) for k in rest_keys})
                self.ready.record(self.side)
                lf = host["lo_f"].to(dev, non_blocking=True)
                lb = host["lo_b"].to(dev, non_blocking=True)
                self.chunk_ready[0].record(self.side)
                chunks.append((0, lf, lb, self.chunk_ready[0]))
            else:
                if use_flow:
                    B = host["lo_f"].shape[0]
                    if B != sum(self.chunk_sizes):
                        raise ValueError("HostPixelStep: batch of the links differs from the batch it was built for")
                    b0 = 0
                    for i, n in enumerate(self.chunk_sizes):
                        lf = host["lo_f"][b0:b0 + n].to(dev, non_blocking=True)
                        lb = host["lo_b"][b0:b0 + n].to(dev, non_blocking=True)
                        self.chunk_ready[i].record(self.side)
                        chunks.append((b0, lf, lb, self.chunk_ready[i]))
                        b0 += n
                t.update({k: host[k].to(dev, non_blocking=True) for k in feat_keys})
                self.ready_feat.record(self.side)
                t.update({k: host[k].to(dev, non_blocking=True) for k in rest_keys})
                self.ready.record(self.side)
        ff = fb = mf = mb = None
        if use_flow and self.sparse:
            (_, lf, lb, ev), = chunks
            main.wait_event(ev)  # main has nothing else to do until the loss; the PPM forward is on `aux`
            if not capturing:
                lf.record_stream(main)
                lb.record_stream(main)
            pair = ops.LazyFlowPair(lf, lb, flow_up=self.flow_up, alpha_1=self.alpha1, alpha_2=self.alpha2)
            (ff, fb), (mf, mb) = pair.flow, pair.mask
        elif use_flow:
            lo_h, lo_w = host["lo_f"].shape[-2:]
            H, W = (8 * lo_h, 8 * lo_w) if self.flow_up else (lo_h, lo_w)
            ff = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
            fb = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
            mf = torch.empty((B, H, W), device=dev, dtype=torch.uint8)
            mb = torch.empty((B, H, W), device=dev, dtype=torch.uint8)
            for b0, lf, lb, ev in chunks:
                main.wait_event(ev)
                if not capturing:
                    lf.record_stream(main)
                    lb.record_stream(main)
                e = b0 + lf.shape[0]
                ops.flow_stage(lf, lb, flow_up=self.flow_up, alpha_1=self.alpha1, alpha_2=self.alpha2,
                               out=(ff[b0:e], fb[b0:e], mf[b0:e], mb[b0:e]))
            mf, mb = mf.view(torch.bool), mb.view(torch.bool)
        # PPM forward on a third stream: it needs only the features, so it runs underneath the tail of
        # the flow kernels instead of after them; the loss (which needs both) joins the two.
        with torch.cuda.stream(self.aux):
            self.aux.wait_event(self.ready_feat)  # also forks aux from the (possibly capturing) main stream
            if not capturing:
                for k in feat_keys:
                    t[k].record_stream(self.aux)
            f12 = torch.cat([t["feat1"], t["feat2"]], dim=0).requires_grad_(True)
            wg = w.detach().requires_grad_(True)
            bg = bias.detach().requires_grad_(True)
            pred12 = ops.featprop(f12, wg, bg, self.gamma, self.clamp, final_norm=True)
        main.wait_event(self.ready)
        main.wait_stream(self.aux)
        if not capturing:
            for v in list(t.values()) + [pred12, f12]:
                v.record_stream(main)
        loss, _, pn, _ = ops.regression_loss_pair(pred12, t["k2"], t["c1"], t["c2"], None, t["k1"], t["c2"], t["c1"],
                                                  self.pos_ratio, flow1=ff, flow2=fb, size=self.size, mask1=mf, mask2=mb)
        loss.backward()
        B = t["feat1"].shape[0]
        self.out["loss"].copy_(loss.detach(), non_blocking=True)
        self.out["pos_num"].copy_(pn, non_blocking=True)
        self.out["d_feat"].copy_(f12.grad.view(2, B, *f12.shape[1:]), non_blocking=True)
        return self.out, (wg.grad, bg.grad)
