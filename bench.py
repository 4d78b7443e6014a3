#!/usr/bin/env python
"""bench.py — the pixel-pretext hot path of PixPro-with-OpticalFlow on B200.

One "step" = one pass of the hot path over one per-GPU batch of synthetic input, forward and
backward:  flow stage (x8 up-sampling fused with chaining of the n_frames-1 links + both
forward-backward consistency masks)  ->  PPM on both views (value_transform 1x1 conv on this
repo's tcgen05 3xTF32 kernel, similarity / relu^2 / propagation / L2-normalise in the sm_100a
kernels)  ->  flow-guided correspondence, positive mask and masked cosine regression loss in both
directions  ->  backward of all of it down to the gradients of the two projector feature maps.

Headline workload = BASELINE.json configs[1]: n_frames=2, 90x160 low-res flow links up-sampled
to 720x1280 (--flow_up, the published setting), batch 64 per GPU, 7x7 grid, 256-d features,
alpha1=0.01 alpha2=0.5 pos_ratio=0.7 p=2 transform_layer=1.  The ResNet-50 backbone is not
part of the path (it stays on cuDNN, BASELINE.json north_star) and is not timed here.
The same invocation also times BASELINE's other single-GPU configurations beside the headline
(`configs`: n6 = n_frames 6 dense chain, g14 = 14x14 grid, g28 = 28x28 grid at batch 32), each
with its own per-kernel roofline, and (`--pretrain`, or the `pretrain` object when the backbone
fits the time budget) the whole DDP pre-training step of main_pretrain.py on synthetic data.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # the CPU arm (oracle port), host cores

Prints ONE JSON line (rank 0).  metric = frames/sec = B * world * n_frames / step_time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pixpro-with-opticalflow_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

H_LO, W_LO = 90, 160
H_FULL, W_FULL = 720, 1280
C_FEAT = 256
ALPHA1, ALPHA2, POS_RATIO, GAMMA, CLAMP = 0.01, 0.5, 0.7, 2.0, 0.0

# BASELINE.json configs[2..4] at their single-GPU shapes, timed beside the headline (configs[1])
EXTRA_CONFIGS = {
    "n6": dict(batch=64, n_frames=6, grid=7, what="BASELINE configs[2]/[4] flow side: 6-frame chained correspondence, dense"),
    "g14": dict(batch=64, n_frames=2, grid=14, what="BASELINE configs[3]/[4]: 14x14 grid (448^2 crops)"),
    "g28": dict(batch=32, n_frames=2, grid=28, what="BASELINE configs[3]: 28x28 grid"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--n-frames", type=int, default=2)
    ap.add_argument("--grid", type=int, default=7)
    ap.add_argument("--cpu-sample", type=int, default=0, help="samples per CPU-arm step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the n6 / g14 / g28 configurations and the torch-GPU baseline")
    ap.add_argument("--no-graph", action="store_true", help="time the eagerly issued step instead of its CUDA-graph replay")
    ap.add_argument("--sparse", action="store_true",
                    help="headline = the sparse-correspondence step (flow stage evaluated only at the loss's grid centres, "
                         "pp_sparse_corr; same loss / counts / gradients, no dense composites or masks)")
    ap.add_argument("--no-overlap", action="store_true", help="PPM forward on the same stream as the flow stage (A/B switch)")
    ap.add_argument("--pretrain", action="store_true",
                    help="time the whole DDP pre-training step (main_pretrain.py --synthetic, BASELINE configs[2]) "
                         "instead of the pixel path alone; metric pretrain frames/s")
    ap.add_argument("--pretrain-batch", type=int, default=128)
    ap.add_argument("--pretrain-frames", type=int, default=6)
    return ap.parse_args()


def workload_name(batch, n_frames, grid):
    return (f"pixel-pretext hot path fwd+bwd: flow stage (x8 upflow+chain of {n_frames - 1} link(s), 2 FB masks, "
            f"{H_LO}x{W_LO}->{H_FULL}x{W_FULL}) + PPM + flow-guided masked regression loss, both directions; "
            f"n_frames={n_frames}, batch {batch}/GPU, {grid}x{grid} grid, {C_FEAT}-d (BASELINE.json configs[1])")


# ------------------------------------------------------------------------------ algorithmic work

def kernel_work(kernel, B, n, G):
    """Algorithmic (bytes, flops) of one STEP's launches of `kernel`: SURVEY.md §8(d) per-sample figures (both
    directions / both views) x the per-GPU batch; flops are USEFUL flops of the unpadded contraction (2*M*N*K),
    not the 3xTF32 tensor work and not the padded tiles.  The caller divides by the launches per step."""
    P = G * G
    C = C_FEAT
    lo = n * 2 * 2 * H_LO * W_LO * 4          # n links, 2 directions, 2 channels
    comp = 2 * 2 * H_FULL * W_FULL * 4        # 2 composite flows
    masks = 2 * H_FULL * W_FULL               # 2 byte masks
    cp = C * P * 4                            # one [C,P] fp32 map
    pp = P * P * 4                            # one [P,P] fp32 map
    ww = C * C * 4
    gemm = 2 * P * P * C                      # one P x P x C contraction
    conv = 2 * C * C * P
    per_sample = {                            # name -> (bytes, flops), both views / directions of ONE sample
        "chain_up": (lo + comp, 0),           # F1: read low-res links, write composites
        "chain_up1": ((lo + comp) // 2, 0),   # F1 of ONE direction (n = 1 fused route: the other composite is written by fb_up)
        # n = 1 fused route, one launch per direction, each reads its low-res link and the opposite composite and writes
        # its mask; the backward one (fb_up_w) also writes its own composite
        "fb_up_w": (lo // 2 + comp + masks // 2, 0),
        "fb_up": (lo // 2 + comp // 2 + masks // 2, 0),
        "chain_dense": ((n + 1) * comp, 0),   # F1': read n dense links, write the composites
        "fb": (comp + masks, 0),              # F2: read composites, write masks
        "fb1": (comp + masks // 2, 0),        # F2 of one direction: its own composite + the gather source, one mask
        "sparse_corr": (lo + 2 * 3 * P * 4, 0),     # at most the links once; writes [3,P] per direction
        "add_flow": (2 * 5 * P * 4, 0),
        "loss_small": (6 * cp, 2 * gemm),     # F3: read q, k, write dq, both directions; one masked contraction each
        "loss_main": (6 * cp, 2 * gemm),
        "ppm_fwd_small": (2 * 3 * cp, 2 * 2 * gemm),      # F4 forward, both views: read feat, val, write out
        "ppm_bwd_small": (2 * 6 * cp, 2 * 3 * gemm),      # F4 backward: read feat, val, out, g, write d_feat, d_val
        "ppm S (tcgen05)": (2 * (cp + pp), 2 * gemm),
        "ppm Y (tcgen05)": (2 * (2 * cp + pp), 2 * gemm),
        "ppm gS (tcgen05)": (2 * (2 * cp + 3 * pp), 2 * gemm),   # useful flops: G once (the TMA route also accumulates G^T: K = 2C)
        "ppm gvh (tcgen05)": (2 * (2 * cp + pp), 2 * gemm),
        "ppm gxh (tcgen05)": (2 * (2 * cp + pp), 2 * gemm),
        "loss M=K*pos^T (tcgen05)": (2 * (2 * cp + P * 8), 2 * gemm),
        "conv1x1 fwd (tcgen05)": (2 * 2 * cp + ww, 2 * conv),
        "conv1x1 dgrad (tcgen05)": (2 * 2 * cp + ww, 2 * conv),
        "conv1x1 wgrad (tcgen05)": (2 * 2 * cp + ww, 2 * conv),
        "conv1x1 wgrad reduce": (0, 0),
        "conv1x1 bias grad": (2 * cp, 0),
        "conv1x1 bias grad reduce": (0, 0),
        "ppm colnorm": (2 * 3 * cp, 0),       # three maps per view are normed per step (feat, val, out)
        "ppm coldiv": (2 * 3 * 2 * cp, 0),
        "ppm colnorm+div": (2 * 3 * 2 * cp, 0),   # norm + scale in one pass: read each map once, write it once (planes not counted)
        "ppm normbwd": (2 * 3 * 3 * cp, 0),
        "loss_dot": (2 * 2 * cp, 0),
        "loss_pos": (2 * pp // 4, 0),
        "loss_centres": (2 * 5 * P * 4, 0),
        "loss_prep": (2 * 5 * P * 4, 0),
        "loss_cnt": (2 * P * 4, 0),
        "loss_final": (0, 0),
        # plane passes of the TMA-fed route (bytes actually moved per step: what each pass must read and write)
        "ppm planes": (2 * (2 * (1 + 5) + (1 + 4)) * cp, 0),      # feat, val: read 1, write out + 4 planes; gy: read 1, write 4 planes
        "conv1x1 planes": (2 * ((1 + 2) * 4) * cp + 3 * ww, 0),   # x (fwd, wgrad), dy (dgrad, wgrad): read 1, write 2 planes each; W^T
        "tc split": (2 * 3 * cp + 3 * ww, 0),                     # k planes of the two loss directions; W planes
    }
    by, fl = per_sample.get(kernel, (0, 0))
    return by * B, fl * B


def step_work(B, n, G, use_flow=True):
    """SURVEY.md §8(d) whole-step algorithmic work: F1+F2 fused bytes, F3+F4 bytes and 28*P^2*C flops per sample."""
    P = G * G
    flow_bytes = (n * 230400 + 14745600 + 1843200) if use_flow else 0
    pix_bytes = (6 + 10) * C_FEAT * P * 4
    return (flow_bytes + pix_bytes) * B, 28 * P * P * C_FEAT * B


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def roofline_of(name, n_launch, ms_total, prof_steps, B, n, G, peaks):
    """Roofline object of one kernel from its serialised event-bracketed device time (ms_total over prof_steps)."""
    by, fl = kernel_work(name, B, n, G)
    per_launch = n_launch / prof_steps
    alg_b, alg_f = by / per_launch, fl / per_launch
    t = ms_total / n_launch * 1e-3
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    gbs = alg_b / t / 1e9 if alg_b else None
    tfs = alg_f / t / 1e12 if alg_f else None
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    tensor = "tcgen05" in name
    out = {"kernel": name, "avg_launch_ms": ms_total / n_launch, "launches_per_step": per_launch,
           "algorithmic_bytes_per_launch": alg_b, "useful_flops_per_launch": alg_f,
           "hbm_gbs": gbs, "hbm_frac": (gbs / hbm_peak) if gbs else None,
           "useful_tflops": tfs, "tensor_frac": (tfs / tf_peak) if tfs else None}
    if tensor:
        out.update({"bound": "tensor", "achieved": tfs, "peak": tf_peak, "unit": "TFLOP/s", "frac": (tfs / tf_peak) if tfs else None,
                    "peak_source": f"{src} bf16_tflops_sustained (cuBLAS bf16; this kernel does 3 TF32 MMAs per useful product, "
                                   "so 1/6 of this peak is its arithmetic ceiling)"})
    else:
        out.update({"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": (gbs / hbm_peak) if gbs else None,
                    "peak_source": f"{src} hbm_gbs"})
    return out


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every few ms
    in a thread (nvidia-smi's own loop is too slow for a 20 ms region); falls back to one
    `nvidia-smi` query per 50 ms if NVML is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None

    def _loop_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        while not self.stop_flag.is_set():
            mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            try:
                mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            self.samples.append((time.perf_counter(), mhz, mask))
            time.sleep(0.002)

    def _loop_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                mask = sum(b for b, v in zip(bits, out[2:6]) if v.strip().lower().startswith("active"))
                self.max_mhz = float(out[1])
                self.samples.append((time.perf_counter(), float(out[0]), mask))
            except Exception:
                pass
            time.sleep(0.05)

    def _loop(self):
        try:
            self._loop_nvml()
        except Exception:
            self._loop_smi()

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        deadline = time.perf_counter() + 5.0  # NVML initialisation can take a while on a fresh box:
        while not self.samples and time.perf_counter() < deadline:  # wait for the first sample before the timed region opens
            time.sleep(0.005)
        self.t0 = time.perf_counter()

    def stop(self):
        t1 = time.perf_counter()
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        inside = [(m, k) for (t, m, k) in self.samples if self.t0 <= t <= t1] or [(m, k) for (_, m, k) in self.samples[-3:]]
        mask = 0
        for _, k in inside:
            mask |= k
        return {"sm_mhz": statistics.median([m for m, _ in inside]) if inside else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for bit, name in self.REASONS.items() if mask & bit), "samples": len(inside)}


# ------------------------------------------------------------------------------ multi-rank bookkeeping

def max_over_ranks(ms, world, device):
    """Step time of the job = the slowest rank's device time."""
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_frames_per_s(batch, world, n_frames, ms_per_step):
    """Whole-job throughput: every rank processes its own `batch` samples per step (weak scaling)."""
    return batch * world * n_frames / (ms_per_step * 1e-3)


# ------------------------------------------------------------------------------ the CUDA arm

def make_inputs(batch, n_frames, grid, seed):
    import torch
    from pixpro_b200 import synth
    n = n_frames - 1
    lo_f, lo_b = synth.flow_fields(batch, max(n, 1), h=H_LO, w=W_LO, seed=seed)
    feat1, feat2, k1, k2 = synth.features(batch, C_FEAT, grid, seed=seed + 1)
    c1 = synth.crop_coords(batch, W_FULL, H_FULL, seed=seed + 2)
    c2 = synth.crop_coords(batch, W_FULL, H_FULL, seed=seed + 3)
    g = torch.Generator().manual_seed(seed + 4)
    w = torch.randn(C_FEAT, C_FEAT, 1, 1, generator=g) / 16.0   # value_transform (transform_layer=1)
    bias = torch.zeros(C_FEAT)
    return dict(lo_f=lo_f, lo_b=lo_b, feat1=feat1, feat2=feat2, k1=k1, k2=k2, c1=c1, c2=c2, w=w, bias=bias)


class Ctx:
    """Per-process state shared by every measured configuration."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.a = a
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the pixel-pretext path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        self.side = torch.cuda.Stream(device=self.dev, priority=-1)
        self.peaks = load_peaks()

    def barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, flush=True):
        """CUDA-event time of each of `steps` calls of fn on the current stream, L2 flushed before each."""
        import torch
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            if flush:
                self.flush_buf.zero_()
            starts[i].record()
            fn()
            stops[i].record()
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in zip(starts, stops)]


def make_hot_path(ctx, batch, n_frames, grid):
    """Returns (hot_path(t, sparse, overlap), device inputs, pinned inputs)."""
    import torch
    from pixpro_b200 import ops
    dev = ctx.dev
    use_flow = n_frames > 1
    host = make_inputs(batch, n_frames, grid, 1234 + ctx.rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    size = (H_FULL, W_FULL)

    def hot_path(t, sparse=False, overlap=True):
        """One pass of the path on device-resident tensors t; returns (loss, pos stats, grads)."""
        w = t["w"].detach().requires_grad_(True)
        bias = t["bias"].detach().requires_grad_(True)
        cur = torch.cuda.current_stream(dev)
        # The PPM forward depends on nothing the flow stage produces.  It is a handful of latency-bound
        # one-block-per-sample launches, so it goes to a HIGH-PRIORITY side stream (its few blocks get the
        # first free SM slots) and runs underneath the HBM-bound flow kernels of the current stream.
        ppm_stream = ctx.side if overlap else cur
        ppm_stream.wait_stream(cur)
        with torch.cuda.stream(ppm_stream):
            # as PixPro.forward does: both views through the PPM as one batch (the leaf is the joint tensor, so the
            # feature gradient arrives in one piece)
            f12 = torch.cat([t["feat1"], t["feat2"]], dim=0).requires_grad_(True)
            pred12 = ops.featprop(f12, w, bias, GAMMA, CLAMP, final_norm=True)  # value transform + PPM, one autograd node
        if use_flow and sparse:
            pair = ops.LazyFlowPair(t["lo_f"], t["lo_b"], flow_up=True, alpha_1=ALPHA1, alpha_2=ALPHA2)
            (ff, fb), (mf, mb) = pair.flow, pair.mask
        elif use_flow:
            ff, fb, mf, mb = ops.flow_stage(t["lo_f"], t["lo_b"], flow_up=True, alpha_1=ALPHA1, alpha_2=ALPHA2)
        else:
            ff = fb = mf = mb = None
        cur.wait_stream(ppm_stream)
        pred12.record_stream(cur)
        f12.record_stream(cur)
        # both loss directions in one launch, on the joint prediction tensor: loss_1 + loss_2 and one gradient tensor
        loss, _, pn, _ = ops.regression_loss_pair(pred12, t["k2"], t["c1"], t["c2"], None, t["k1"], t["c2"], t["c1"], POS_RATIO,
                                                  flow1=ff, flow2=fb, size=size, mask1=mf, mask2=mb)
        pn1, pn2 = pn[0], pn[1]
        loss.backward()
        B = t["feat1"].shape[0]
        return loss.detach(), pn1, pn2, f12.grad[:B], f12.grad[B:]

    return hot_path, d, pinned


def graphed(ctx, fn):
    """fn captured once into a CUDA graph; returns the replay callable (fn itself with --no-graph)."""
    import torch
    if ctx.a.no_graph:
        return fn
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = fn()
    def replay(graph=graph, keep=keep):  # the closure keeps the graph and its output buffers alive
        graph.replay()

    for _ in range(2):
        replay()
    return replay


def profile_kernels(ctx, hot_path, d, batch, n_frames, grid, sparse, steps):
    """Per-kernel device times from the library's own per-launch event brackets in a SERIALISED eager pass: one
    stream (no PPM side stream, no side streams inside ops), so a kernel's time is its own, not its overlap partner's."""
    from pixpro_b200 import _cabi, ops
    prof_steps = max(3, min(steps, 10))
    ops.set_serial(True)
    try:
        for _ in range(2):
            hot_path(d, sparse=sparse, overlap=False)
        _cabi.profile_enable(True)
        ctx.timed(lambda: hot_path(d, sparse=sparse, overlap=False), prof_steps)
        rep = _cabi.profile_report()
        _cabi.profile_enable(False)
    finally:
        ops.set_serial(False)
    tot = sum(ms for _, ms in rep.values()) or 1.0
    kernels = {k: {"launches_per_step": nl / prof_steps, "ms_per_step": ms / prof_steps, "share": ms / tot}
               for k, (nl, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
    n = n_frames - 1
    roofs = {k: roofline_of(k, nl, ms, prof_steps, batch, n, grid, ctx.peaks) for k, (nl, ms) in rep.items()}
    top = max(rep, key=lambda k: rep[k][1]) if rep else None
    tcs = [k for k in rep if "tcgen05" in k and roofs[k]["useful_flops_per_launch"]]
    top_tc = max(tcs, key=lambda k: rep[k][1]) if tcs else None
    for k, r in roofs.items():   # compact per-kernel fractions beside the times
        kernels[k]["hbm_frac"] = r["hbm_frac"]
        if r["tensor_frac"] is not None:
            kernels[k]["tensor_frac"] = r["tensor_frac"]
    return kernels, (roofs[top] if top else None), (roofs[top_tc] if top_tc else None), tot / prof_steps


def dram_traffic(kernel, batch, n):
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            return json.load(f).get(f"{kernel}:B{batch}:n{n}")
    except (OSError, ValueError):
        return None


def measure_config(ctx, batch, n_frames, grid, steps, warmup, sparse=False, e2e=True, sparse_beside=True):
    """Times one configuration: device-resident step (graph replay), optional e2e from pinned host buffers through
    PinnedFlowStager -> HostPixelStep, optional sparse-correspondence step beside it, serialised per-kernel pass."""
    import torch
    from pixpro_b200 import _cabi
    a = ctx.a
    world, dev = ctx.world, ctx.dev
    use_flow = n_frames > 1
    hot_path, d, pinned = make_hot_path(ctx, batch, n_frames, grid)
    overlap = not a.no_overlap
    res = {}

    # -------- device-resident throughput ("value") --------
    for _ in range(warmup):
        hot_path(d, sparse=sparse, overlap=overlap)
    # The step is ~0.5-3 ms of device work issued by ~30 host-side calls: replayed from ONE CUDA graph so that
    # the number measures the kernels, not the host's launch rate (the graph holds exactly the launches the
    # eager step makes; --no-graph times the eager step).
    n0 = _cabi.launch_count()
    hot_path(d, sparse=sparse, overlap=overlap)
    launches_per_step = _cabi.launch_count() - n0
    step_fn = graphed(ctx, lambda: hot_path(d, sparse=sparse, overlap=overlap))
    sampler = ClockSampler(ctx.local)
    ctx.barrier()
    if ctx.rank == 0:
        sampler.start()
    per_step = ctx.timed(step_fn, steps)
    ctx.barrier()
    res["clocks"] = sampler.stop() if ctx.rank == 0 else None
    total_ms = max_over_ranks(sum(per_step), world, dev)
    ms = total_ms / steps
    res.update({"ms_per_step": ms, "us_per_batch": ms * 1e3, "value": aggregate_frames_per_s(batch, world, n_frames, ms),
                "unit": "frames/s", "per_gpu_batch": batch, "n_frames": n_frames, "grid": grid,
                "gpu_launches_per_step": launches_per_step, "timed_region_ms": total_ms,
                "ms_per_step_median": statistics.median(per_step), "ms_per_step_min": min(per_step)})
    sb, sf = step_work(batch, n_frames - 1, grid, use_flow)
    res["step_roofline"] = {"algorithmic_bytes": sb, "useful_flops": sf, "hbm_gbs": sb / (ms * 1e-3) / 1e9,
                            "hbm_frac": sb / (ms * 1e-3) / 1e9 / float(ctx.peaks.get("hbm_gbs", 6650.0)),
                            "useful_tflops": sf / (ms * 1e-3) / 1e12,
                            "tensor_frac": sf / (ms * 1e-3) / 1e12 / float(ctx.peaks.get("bf16_tflops_sustained", 1400.0)),
                            "note": "SURVEY.md 8(d) fused figures per sample (F1+F2 bytes, F3+F4 bytes, 28*P^2*C flops) x batch / step time"}

    # -------- end to end from pinned host buffers ("e2e") --------
    # through the package's host-side entries: the loader-side PinnedFlowStager collates per-sample link slices into
    # its pinned [B,n,2,h,w] batch, HostPixelStep copies links, crop descriptors, features and keys from pinned host
    # memory, runs the path and returns loss, positive counts and feature gradients to pinned host memory,
    # synchronising before it returns.
    if e2e:
        from pixpro_b200.flowstore import PinnedFlowStager
        from pixpro_b200.host_step import HostPixelStep
        e2e_keys = ["feat1", "feat2", "k1", "k2", "c1", "c2"]
        host_in = {k: pinned[k] for k in e2e_keys}
        if use_flow:
            stager = PinnedFlowStager(batch, n_frames - 1, H_LO, W_LO, buffers=1)   # one buffer: the graph's fixed endpoint
            samples = [(pinned["lo_f"][b], pinned["lo_b"][b]) for b in range(batch)]
            host_in["lo_f"], host_in["lo_b"] = stager.collate(samples)
        hstep = HostPixelStep(dev, batch, C_FEAT, grid, size=(H_FULL, W_FULL), gamma=GAMMA, clamp=CLAMP, pos_ratio=POS_RATIO,
                              alpha1=ALPHA1, alpha2=ALPHA2, sparse=sparse)
        for _ in range(max(3, warmup // 2)):
            hstep(host_in, d["w"], d["bias"])
        ctx.barrier()
        e2e_ms = max_over_ranks(sum(ctx.timed(lambda: hstep(host_in, d["w"], d["bias"]), steps)), world, dev) / steps
        ctx.barrier()
        res["e2e"] = {"value": aggregate_frames_per_s(batch, world, n_frames, e2e_ms), "unit": "frames/s", "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": hstep.h2d_bytes(host_in), "d2h_bytes_per_step": hstep.d2h_bytes(),
                      "entry": "pixpro_b200.flowstore.PinnedFlowStager -> pixpro_b200.host_step.HostPixelStep"}

    # -------- the same step through the sparse correspondence path (reported beside the headline) --------
    if use_flow and not sparse and sparse_beside:
        for _ in range(3):
            hot_path(d, sparse=True, overlap=overlap)
        ref_out = hot_path(d, sparse=False, overlap=overlap)
        sp_out = hot_path(d, sparse=True, overlap=overlap)
        identical = all(bool(torch.equal(x, y)) for x, y in zip(ref_out, sp_out))
        sp_fn = graphed(ctx, lambda: hot_path(d, sparse=True, overlap=overlap))
        ctx.barrier()
        sp_ms = max_over_ranks(sum(ctx.timed(sp_fn, steps)), world, dev) / steps
        res["sparse_correspondence"] = {
            "what": "same step with the flow stage evaluated only at the loss's grid centres (pp_sparse_corr): "
                    "no dense composites / FB masks; outputs compared with the dense step",
            "ms_per_step": sp_ms, "value": aggregate_frames_per_s(batch, world, n_frames, sp_ms), "unit": "frames/s",
            "outputs_bit_identical_to_dense_step": identical}

    # -------- per-kernel device times -> rooflines --------
    kernels, roof, roof_tc, ser_ms = profile_kernels(ctx, hot_path, d, batch, n_frames, grid, sparse, steps)
    if roof is not None:
        roof["traffic"] = dram_traffic(roof["kernel"], batch, n_frames - 1)
    res["kernels"] = kernels
    res["roofline"] = roof
    if roof_tc is not None and (roof is None or roof_tc["kernel"] != roof["kernel"]):
        res["roofline_tensor"] = roof_tc
    res["kernel_ms_serialised"] = ser_ms
    return res, (hot_path, d)


def torch_gpu_baseline(ctx, batch, n_frames, grid, steps):
    """The reference's torch-op sequence for the same path (oracle/torch_restatement.py: the ops contrast/util.py and
    contrast/models/PixPro.py issue, in their order) run eagerly on the same B200 — what a user of the unmodified
    reference gets on this GPU for this path.  A baseline leg, like cpu_baseline: the product never calls it."""
    import torch
    import torch.nn.functional as F
    from oracle import torch_restatement as TR
    dev = ctx.dev
    t = {k: v.to(dev) for k, v in make_inputs(batch, n_frames, grid, 1234 + ctx.rank).items()}
    size = (H_FULL, W_FULL)

    def step():
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=False):
            ff, fb, mf, mb = TR.torch_flow_stage(t["lo_f"], t["lo_b"], ALPHA1, ALPHA2)
        f1 = t["feat1"].detach().requires_grad_(True)
        f2 = t["feat2"].detach().requires_grad_(True)
        w = t["w"].detach().requires_grad_(True)
        b = t["bias"].detach().requires_grad_(True)
        tot = 0
        for f, k, cq, ck, fl, mk in ((f1, t["k2"], t["c1"], t["c2"], ff, mf), (f2, t["k1"], t["c2"], t["c1"], fb, mb)):
            pred = F.normalize(TR.torch_featprop_fn(f, F.conv2d(f, w, b), GAMMA, CLAMP), dim=1)
            l, _ = TR.torch_regression_loss(pred, k, cq, ck, POS_RATIO, flow=fl, size=size, mask=mk)
            tot = tot + l
        tot.backward()
        return tot.detach()

    reps = max(3, min(steps, 5))
    for _ in range(2):
        step()
    ms = sum(ctx.timed(step, reps)) / reps
    return {"value": batch * n_frames / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "steps": reps,
            "what": "eager PyTorch restatement of the reference's op sequence (oracle/torch_restatement.py) on the same GPU, "
                    "fp32, cuDNN sampler off (ATen native kernels), inputs resident in HBM; one rank"}


def run_b200(a):
    import torch
    import torch.distributed as dist
    ctx = Ctx(a)
    if a.pretrain:
        return run_pretrain(ctx)
    res, _ = measure_config(ctx, a.batch, a.n_frames, a.grid, a.steps, a.warmup, sparse=a.sparse)
    default_workload = (a.batch, a.n_frames, a.grid) == (64, 2, 7) and not a.sparse
    extra = {}
    if not a.no_extra and default_workload:
        xs = max(10, a.steps // 2)
        for name, c in EXTRA_CONFIGS.items():
            try:
                r, _ = measure_config(ctx, c["batch"], c["n_frames"], c["grid"], xs, max(3, a.warmup), e2e=False, sparse_beside=False)
                r["what"] = c["what"]
                r["steps"] = xs
                r.pop("clocks", None)
                extra[name] = r
            except Exception as e:  # a failing side configuration must not take the headline down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
    # the whole DDP pre-training step (BASELINE configs[2]) beside the pixel path, so that the driver's 1/2/4/8-GPU scaling
    # runs carry it: every rank takes part (NCCL gradient all-reduce), a few steps only
    pre = None
    if not a.no_extra and default_workload:
        try:
            pre = measure_pretrain(ctx, steps=8, warmup=6)   # warm-up covers the momentum branch's graph capture (after 3 eager steps)
            for k in ("higher_is_better", "scaling", "vs_baseline", "data"):
                pre.pop(k, None)
        except Exception as e:
            pre = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    tgb = None
    if not a.no_extra and ctx.rank == 0:
        try:
            tgb = torch_gpu_baseline(ctx, a.batch, a.n_frames, a.grid, a.steps)
        except Exception as e:
            tgb = {"error": f"{type(e).__name__}: {e}"}
    if ctx.world > 1:
        dist.barrier()
    if ctx.rank != 0:
        if dist.is_initialized():
            dist.destroy_process_group()
        return
    line = {
        "metric": "pixel-pretext hot path (flow chain + FB mask + PPM + flow-guided loss, fwd+bwd) frames/sec",
        "value": res["value"], "unit": "frames/s", "n_gpus": ctx.world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": res["ms_per_step"], "us_per_batch": res["us_per_batch"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a.batch, a.n_frames, a.grid), "per_gpu_batch": a.batch, "n_frames": a.n_frames,
                   "grid": a.grid, "flow_up": True, "l2": "explicit flush (256 MiB memset) before every timed step",
                   "issue": "eager" if a.no_graph else "one CUDA graph replay per step (the eager step's launches, captured once)",
                   "sharding": "independent samples per rank, no data-path collective"},
        "e2e": res.get("e2e"),
        "gpu_launches": res["gpu_launches_per_step"] * a.steps,
        "gpu_launches_per_step": res["gpu_launches_per_step"],
        "clocks": res["clocks"],
        "roofline": res["roofline"],
        "step_roofline": res["step_roofline"],
        "kernels": res["kernels"],
        "samples_per_s": res["value"] / a.n_frames,
        "timed_region_ms": res["timed_region_ms"],
    }
    if "roofline_tensor" in res:
        line["roofline_tensor"] = res["roofline_tensor"]
    if "sparse_correspondence" in res:
        line["sparse_correspondence"] = res["sparse_correspondence"]
    if a.sparse:
        line["config"]["flow_stage"] = "sparse correspondence (pp_sparse_corr): evaluated at the loss's grid centres only"
    if extra:
        line["configs"] = extra
    if pre is not None:
        line["pretrain_ddp"] = pre
    if tgb is not None:
        line["torch_gpu_baseline"] = tgb
    if not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_arm(a.batch, a.n_frames, a.grid, steps=8, warmup=1, budget_s=25.0, sample=a.cpu_sample)
    emit(line)
    if dist.is_initialized():
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ the DDP pre-training step

def measure_pretrain(ctx, steps, warmup):
    """The whole pre-training step of main_pretrain.py (drop-in contrast.models.PixPro around an eager PyTorch ResNet-50 +
    DDP gradient all-reduce over NCCL + LARS + this repo's flow stage), synthetic data, BASELINE configs[2] shape.
    Collective: every rank calls it.  frames/s = B * world * n_frames / step time (max over ranks); rank 0 gets the record."""
    import torch
    import torch.distributed as dist
    import main_pretrain as MP
    a = ctx.a
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=ctx.dev)
    opt = MP.synthetic_options(batch_size=a.pretrain_batch, n_frames=a.pretrain_frames, amp="bf16")
    trainer = MP.SyntheticTrainer(opt, ctx.dev)
    warmup = max(3, warmup)
    for _ in range(warmup):
        trainer.step()
    sampler = ClockSampler(ctx.local)
    ctx.barrier()
    if ctx.rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        trainer.step()
    e1.record()
    ctx.barrier()
    clocks = sampler.stop() if ctx.rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1), ctx.world, ctx.dev) / steps
    fps = a.pretrain_batch * ctx.world * a.pretrain_frames / (ms * 1e-3)
    rec = {"metric": "PixPro+OF pretrain frames/sec (whole DDP step: ResNet-50 + projector on cuDNN/PyTorch, pixel path + "
                     "EMA + LARS on this repo's kernels, NCCL gradient all-reduce)",
           "value": fps, "unit": "frames/s", "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": f"main_pretrain.py --synthetic: PixPro+OF ResNet-50, n_frames={a.pretrain_frames}, "
                                  f"batch {a.pretrain_batch}/GPU, 224^2 crops, 7x7 grid, bf16 autocast, DDP (BASELINE.json configs[2])",
                      "per_gpu_batch": a.pretrain_batch, "n_frames": a.pretrain_frames, "flow_stage": trainer.flow_mode},
           "samples_per_s": fps / a.pretrain_frames, "clocks": clocks, "gpu_launches": trainer.launches(),
           "step_breakdown_ms": trainer.breakdown()}
    del trainer
    torch.cuda.empty_cache()
    return rec


def run_pretrain(ctx):
    """`--pretrain`: the DDP pre-training step as the headline line."""
    import torch.distributed as dist
    rec = measure_pretrain(ctx, ctx.a.steps, ctx.a.warmup)
    if ctx.rank == 0:
        emit(rec)
    if dist.is_initialized():
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ the CPU arm

def cpu_arm(batch, n_frames, grid, steps, warmup, budget_s, sample=0):
    """The oracle port of the reference's path (oracle/pixpro_oracle.c, OpenMP over all host cores) on a bounded
    sample of the same workload: `steps` timed steps after `warmup`, each over `bs` samples, bs chosen from one
    calibration step so that the whole run fits `budget_s` seconds."""
    import numpy as np
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    use_flow = n_frames > 1

    def build(bs):
        t = {k: v.numpy() for k, v in make_inputs(bs, n_frames, grid, 1234).items()}
        w2 = t["w"][:, :, 0, 0].astype(np.float64)

        def conv(x):
            return (np.einsum("oc,bchw->bohw", w2, x.astype(np.float64)) + t["bias"][None, :, None, None]).astype(np.float32)

        def step():
            ff = fb = mf = mb = None
            if use_flow:
                ff, fb, mf, mb = orc.flow_stage(t["lo_f"], t["lo_b"], flow_up=True, alpha_1=ALPHA1, alpha_2=ALPHA2)
            tot = 0.0
            for feat, key, cq, ck, fl, mk in ((t["feat1"], t["k2"], t["c1"], t["c2"], ff, mf),
                                              (t["feat2"], t["k1"], t["c2"], t["c1"], fb, mb)):
                val = conv(feat)
                pred = orc.featprop(feat, val, GAMMA, CLAMP, True)
                o = orc.regression_loss(pred, key, cq, ck, POS_RATIO, flow=fl, size=(H_FULL, W_FULL), mask=mk)
                dfs, dv = orc.featprop_bwd(feat, val, o["dq"], GAMMA, CLAMP, True)
                _ = dfs + np.einsum("oc,bohw->bchw", w2, dv.astype(np.float64)).astype(np.float32)
                tot += o["loss"]
            return tot
        return step

    bs = sample or batch
    if not sample:
        cal_bs = min(batch, 4)
        cal = build(cal_bs)
        cal()
        t0 = time.perf_counter()
        cal()
        per_sample = (time.perf_counter() - t0) / cal_bs
        bs = int(max(1, min(batch, budget_s / max(steps + warmup, 1) / max(per_sample, 1e-6))))
    step = build(bs)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": bs * n_frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{bs} of {batch} samples per step, {steps} steps after {warmup} warm-up, {dt * 1e3:.0f} ms/step",
            "ms_per_step": dt * 1e3, "steps": steps, "warmup": warmup,
            "note": "reference is pure Python/PyTorch and cannot travel to the GPU box; the C port (pinned bit-exact "
                    "against it, oracle/pin_against_reference.py) is the CPU arm"}


def run_reference(a):
    """`--impl reference`: rank 0 times the CPU arm with exactly --steps / --warmup steps (each a bounded sample of
    the workload, sized so the run ends within ~3 minutes); other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    base = cpu_arm(a.batch, a.n_frames, a.grid, steps=a.steps, warmup=a.warmup, budget_s=150.0, sample=a.cpu_sample)
    line = {
        "impl": "reference",
        "metric": "pixel-pretext hot path (flow chain + FB mask + PPM + flow-guided loss, fwd+bwd) frames/sec",
        "value": base["value"], "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", a.gpus)),
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a.batch, a.n_frames, a.grid), "per_gpu_batch": a.batch, "n_frames": a.n_frames,
                   "grid": a.grid, "flow_up": True, "l2": "explicit flush (256 MiB memset) before every timed step",
                   "issue": "eager" if a.no_graph else "one CUDA graph replay per step (the eager step's launches, captured once)",
                   "sharding": "independent samples per rank, no data-path collective"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_JSON_OUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    # Libraries write banners to fd 1 (NCCL prints its version line there when a communicator is created):
    # keep the original stdout for the JSON line only and send everything else to stderr.
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
