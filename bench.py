#!/usr/bin/env python
"""bench.py — the pixel-pretext hot path of PixPro-with-OpticalFlow on B200.

One "step" = one pass of the hot path over one per-GPU batch of synthetic input, forward and
backward:  flow stage (x8 up-sampling fused with chaining of the n_frames-1 links + both
forward-backward consistency masks)  ->  PPM on both views (value_transform 1x1 conv on
cuDNN, similarity / relu^2 / propagation / L2-normalise in the sm_100a kernels)  ->  flow-guided
correspondence, positive mask and masked cosine regression loss in both directions  ->
backward of all of it down to the gradients of the two projector feature maps.

Default workload = BASELINE.json configs[1]: n_frames=2, 90x160 low-res flow links up-sampled
to 720x1280 (--flow_up, the published setting), batch 64 per GPU, 7x7 grid, 256-d features,
alpha1=0.01 alpha2=0.5 pos_ratio=0.7 p=2 transform_layer=1.  The ResNet-50 backbone is not
part of the path (it stays on cuDNN, BASELINE.json north_star) and is not timed here.

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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--n-frames", type=int, default=2)
    ap.add_argument("--grid", type=int, default=7)
    ap.add_argument("--cpu-sample", type=int, default=0, help="samples per CPU-arm step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eagerly issued step instead of its CUDA-graph replay")
    ap.add_argument("--sparse", action="store_true",
                    help="headline = the sparse-correspondence step (flow stage evaluated only at the loss's grid centres, "
                         "pp_sparse_corr; same loss / counts / gradients, no dense composites or masks)")
    ap.add_argument("--no-overlap", action="store_true", help="PPM forward on the same stream as the flow stage (A/B switch)")
    return ap.parse_args()


def workload_name(a):
    return (f"pixel-pretext hot path fwd+bwd: flow stage (x8 upflow+chain of {a.n_frames - 1} link(s), 2 FB masks, "
            f"{H_LO}x{W_LO}->{H_FULL}x{W_FULL}) + PPM + flow-guided masked regression loss, both directions; "
            f"n_frames={a.n_frames}, batch {a.batch}/GPU, {a.grid}x{a.grid} grid, {C_FEAT}-d (BASELINE.json configs[1])")


def algorithmic_bytes(kernel, B, n, G=7):
    """Algorithmic bytes of one STEP's launches of `kernel` (SURVEY.md §8(d) per-sample figures, both
    directions / both views, x the per-GPU batch); the caller divides by the launches per step."""
    lo = n * 2 * 2 * H_LO * W_LO * 4          # n links, 2 directions, 2 channels
    comp = 2 * 2 * H_FULL * W_FULL * 4        # 2 composite flows
    masks = 2 * H_FULL * W_FULL               # 2 byte masks
    cp4 = C_FEAT * G * G * 4                  # one [C,P] fp32 map
    per_sample = {
        "chain_up": lo + comp,                # F1: read low-res links, write composites
        "chain_dense": (n + 1) * comp,        # F1': read n dense links, write the composites
        "fb": comp + masks,                   # F2: read composites, write masks
        "loss_small": 6 * cp4,                # F3: read q, k, write dq, both directions
        "ppm_fwd_small": 2 * 3 * cp4,         # F4 forward, both views: read feat, val, write out
        "ppm_bwd_small": 2 * 6 * cp4,         # F4 backward: read feat, val, out, g, write d_feat, d_val
    }
    return per_sample.get(kernel, 0) * B


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

def make_inputs(a, seed):
    import torch
    from pixpro_b200 import synth
    n = a.n_frames - 1
    lo_f, lo_b = synth.flow_fields(a.batch, max(n, 1), h=H_LO, w=W_LO, seed=seed)
    feat1, feat2, k1, k2 = synth.features(a.batch, C_FEAT, a.grid, seed=seed + 1)
    c1 = synth.crop_coords(a.batch, W_FULL, H_FULL, seed=seed + 2)
    c2 = synth.crop_coords(a.batch, W_FULL, H_FULL, seed=seed + 3)
    g = torch.Generator().manual_seed(seed + 4)
    w = torch.randn(C_FEAT, C_FEAT, 1, 1, generator=g) / 16.0   # value_transform (transform_layer=1)
    bias = torch.zeros(C_FEAT)
    return dict(lo_f=lo_f, lo_b=lo_b, feat1=feat1, feat2=feat2, k1=k1, k2=k2, c1=c1, c2=c2, w=w, bias=bias)


def run_b200(a):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from pixpro_b200 import _cabi, ops

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the pixel-pretext path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False          # keep the 1x1 value transform in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    use_flow = a.n_frames > 1
    host = make_inputs(a, 1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    size = (H_FULL, W_FULL)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    side = torch.cuda.Stream(device=dev, priority=-1)

    def hot_path(t, sparse=a.sparse):
        """One pass of the path on device-resident tensors t; returns (loss, pos stats, grads)."""
        f1 = t["feat1"].detach().requires_grad_(True)
        f2 = t["feat2"].detach().requires_grad_(True)
        w = t["w"].detach().requires_grad_(True)
        bias = t["bias"].detach().requires_grad_(True)
        cur = torch.cuda.current_stream(dev)
        # The PPM forward depends on nothing the flow stage produces.  It is a handful of latency-bound
        # one-block-per-sample launches, so it goes to a HIGH-PRIORITY side stream (its few blocks get the
        # first free SM slots) and runs underneath the HBM-bound flow kernels of the current stream.
        ppm_stream = cur if a.no_overlap else side
        ppm_stream.wait_stream(cur)
        with torch.cuda.stream(ppm_stream):
            # as PixPro.forward does: both views through the PPM as one batch
            f12 = torch.cat([f1, f2], dim=0)
            pred12 = ops.ppm(f12, ops.conv1x1(f12, w, bias), GAMMA, CLAMP, final_norm=True)
        if use_flow and sparse:
            pair = ops.LazyFlowPair(t["lo_f"], t["lo_b"], flow_up=True, alpha_1=ALPHA1, alpha_2=ALPHA2)
            (ff, fb), (mf, mb) = pair.flow, pair.mask
        elif use_flow:
            ff, fb, mf, mb = ops.flow_stage(t["lo_f"], t["lo_b"], flow_up=True, alpha_1=ALPHA1, alpha_2=ALPHA2)
        else:
            ff = fb = mf = mb = None
        cur.wait_stream(ppm_stream)
        pred12.record_stream(cur)
        pred1, pred2 = pred12.chunk(2, dim=0)
        # both loss directions in one launch
        l12, pn, _ = ops.regression_loss_pair(pred1, t["k2"], t["c1"], t["c2"], pred2, t["k1"], t["c2"], t["c1"], POS_RATIO,
                                              flow1=ff, flow2=fb, size=size, mask1=mf, mask2=mb)
        loss = l12[0] + l12[1]
        pn1, pn2 = pn[0], pn[1]
        loss.backward()
        return loss.detach(), pn1, pn2, f1.grad, f2.grad

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush=True):
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            if flush:
                flush_buf.zero_()
            starts[i].record()
            fn()
            stops[i].record()
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in zip(starts, stops)]

    # -------- device-resident throughput ("value") --------
    for _ in range(a.warmup):
        hot_path(d)
    # The step is ~0.8 ms of device work issued by ~30 host-side calls: replayed from ONE CUDA graph so that
    # the number measures the kernels, not the host's launch rate (the graph holds exactly the launches the
    # eager step makes; --no-graph times the eager step).
    n0 = _cabi.launch_count()
    hot_path(d)
    launches_per_step = _cabi.launch_count() - n0
    step_fn = lambda: hot_path(d)
    if not a.no_graph:
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            graph_out = hot_path(d)
        step_fn = graph.replay
        for _ in range(2):
            step_fn()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    per_step = timed(step_fn, a.steps)
    launches = launches_per_step * a.steps
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(sum(per_step), world, dev)
    ms_per_step = total_ms / a.steps
    frames_per_s = aggregate_frames_per_s(a.batch, world, a.n_frames, ms_per_step)

    # -------- end to end from pinned host buffers ("e2e") --------
    # through the package's host-buffer entry (pixpro_b200.host_step.HostPixelStep): every step copies
    # links, crop descriptors, features and keys from pinned host memory and returns loss, positive
    # counts and the feature gradients to pinned host memory, synchronising before it returns.
    from pixpro_b200.host_step import HostPixelStep
    e2e_keys = ["feat1", "feat2", "k1", "k2", "c1", "c2"] + (["lo_f", "lo_b"] if use_flow else [])
    host_in = {k: pinned[k] for k in e2e_keys}
    hstep = HostPixelStep(dev, a.batch, C_FEAT, a.grid, size=size, gamma=GAMMA, clamp=CLAMP, pos_ratio=POS_RATIO,
                          alpha1=ALPHA1, alpha2=ALPHA2, sparse=a.sparse)
    h2d = hstep.h2d_bytes(host_in)
    d2h = hstep.d2h_bytes()

    def e2e_step():
        hstep(host_in, d["w"], d["bias"])

    for _ in range(max(3, a.warmup // 2)):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks(sum(timed(e2e_step, a.steps)), world, dev) / a.steps
    barrier()
    e2e_fps = aggregate_frames_per_s(a.batch, world, a.n_frames, e2e_ms)

    # -------- the same step through the sparse correspondence path (reported beside the headline) --------
    sparse_info = None
    if use_flow and not a.sparse:
        for _ in range(3):
            hot_path(d, sparse=True)
        n0 = _cabi.launch_count()
        ref_out = hot_path(d, sparse=False)
        sp_out = hot_path(d, sparse=True)
        sp_launches = (_cabi.launch_count() - n0) // 2  # not used for the headline count
        identical = all(bool(torch.equal(x, y)) for x, y in zip(ref_out, sp_out))
        sp_fn = lambda: hot_path(d, sparse=True)
        if not a.no_graph:
            torch.cuda.synchronize()
            sp_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(sp_graph):
                sp_graph_out = hot_path(d, sparse=True)
            sp_fn = sp_graph.replay
            for _ in range(2):
                sp_fn()
        barrier()
        sp_ms = max_over_ranks(sum(timed(sp_fn, a.steps)), world, dev) / a.steps
        sp_hstep = HostPixelStep(dev, a.batch, C_FEAT, a.grid, size=size, gamma=GAMMA, clamp=CLAMP, pos_ratio=POS_RATIO,
                                 alpha1=ALPHA1, alpha2=ALPHA2, sparse=True)
        for _ in range(max(3, a.warmup // 2)):
            sp_hstep(host_in, d["w"], d["bias"])
        barrier()
        sp_e2e_ms = max_over_ranks(sum(timed(lambda: sp_hstep(host_in, d["w"], d["bias"]), a.steps)), world, dev) / a.steps
        barrier()
        sparse_info = {"what": "same step with the flow stage evaluated only at the loss's grid centres (pp_sparse_corr): "
                               "no dense composites / FB masks; outputs compared with the dense step below",
                       "ms_per_step": sp_ms, "value": aggregate_frames_per_s(a.batch, world, a.n_frames, sp_ms),
                       "e2e_ms_per_step": sp_e2e_ms, "e2e_value": aggregate_frames_per_s(a.batch, world, a.n_frames, sp_e2e_ms),
                       "unit": "frames/s", "outputs_bit_identical_to_dense_step": identical}
        del sp_launches

    # -------- per-kernel device times -> roofline of the dominant kernel --------
    _cabi.profile_enable(True)
    prof_steps = min(a.steps, 10)
    timed(lambda: hot_path(d), prof_steps)
    rep = _cabi.profile_report()
    _cabi.profile_enable(False)
    tot_kernel_ms = sum(ms for _, ms in rep.values()) or 1.0
    kernels = {k: {"launches_per_step": n / prof_steps, "ms_per_step": ms / prof_steps, "share": ms / tot_kernel_ms}
               for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
    roofline = None
    if rep:
        top = max(rep, key=lambda k: rep[k][1])
        n_l, ms = rep[top]
        alg = algorithmic_bytes(top, a.batch, a.n_frames - 1, a.grid) / (n_l / prof_steps)  # per launch
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg / (ms / n_l * 1e-3) / 1e9 if alg else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(
                f"{top}:B{a.batch}:n{a.n_frames - 1}")
        except (OSError, ValueError):
            pass
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "algorithmic_bytes_per_launch": alg, "avg_launch_ms": ms / n_l,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": "pixel-pretext hot path (flow chain + FB mask + PPM + flow-guided loss, fwd+bwd) frames/sec",
        "value": frames_per_s, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "us_per_batch": ms_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "per_gpu_batch": a.batch, "n_frames": a.n_frames, "grid": a.grid,
                   "flow_up": True, "l2": "explicit flush (256 MiB memset) before every timed step",
                   "issue": "eager" if a.no_graph else "one CUDA graph replay per step (the eager step's launches, captured once)",
                   "sharding": "independent samples per rank, no data-path collective"},
        "e2e": {"value": e2e_fps, "unit": "frames/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "gpu_launches_per_step": launches / a.steps,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "samples_per_s": frames_per_s / a.n_frames,
    }
    if sparse_info is not None:
        line["sparse_correspondence"] = sparse_info
    if a.sparse:
        line["config"]["flow_stage"] = "sparse correspondence (pp_sparse_corr): evaluated at the loss's grid centres only"
    if not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_arm(a, steps=12, warmup=1)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ the CPU arm

def cpu_arm(a, steps, warmup):
    """The oracle port of the reference's path (oracle/pixpro_oracle.c, OpenMP over all host
    cores) on a bounded sample of the same workload."""
    import numpy as np
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    bs = a.cpu_sample or a.batch
    sub = argparse.Namespace(**vars(a))
    sub.batch = bs
    t = {k: v.numpy() for k, v in make_inputs(sub, 1234).items()}
    use_flow = a.n_frames > 1
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

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": bs * a.n_frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{bs} of {a.batch} samples per step, {steps} steps, {dt * 1e3:.0f} ms/step",
            "ms_per_step": dt * 1e3,
            "note": "reference is pure Python/PyTorch and cannot travel to the GPU box; the C port (pinned bit-exact "
                    "against it) is the CPU arm"}


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    base = cpu_arm(a, steps=max(1, min(a.steps, 5)), warmup=max(1, min(a.warmup, 2)))
    line = {
        "impl": "reference",
        "metric": "pixel-pretext hot path (flow chain + FB mask + PPM + flow-guided loss, fwd+bwd) frames/sec",
        "value": base["value"], "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", a.gpus)),
        "steps": max(1, min(a.steps, 5)), "warmup": max(1, min(a.warmup, 2)), "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "per_gpu_batch": a.batch, "n_frames": a.n_frames, "grid": a.grid,
                   "flow_up": True},
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
