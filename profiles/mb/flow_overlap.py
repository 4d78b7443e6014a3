"""Flow stage at the bench size (B=64, n links, 90x160 -> 720x1280) scheduled three ways:
  whole   one pp_flow_stage call (chain_up of all samples, then fb of all samples)
  chunkS  sample chunks of S on ONE stream (the masks read their composites from L2)
  overS   sample chunks of S, the masks of chunk i on a second stream beside the chain of chunk i+1
Device time per stage (CUDA events, L2 flushed before each repetition) and the mask checksum."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import ops, synth  # noqa: E402

def flow_stage_chunked(lo_fwd, lo_bwd, chunk, side=None, done=None, flow_up=True, alpha_1=0.01, alpha_2=0.5, out=None):
    """(experiment, measured NOT to help: profiles/r02_u_flow_overlap.txt) flow_stage() over sample chunks: the composites of `chunk` samples (2 x chunk x 7.4 MB at 720x1280) are still
    in L2 when their FB masks are computed, and with `side` (a torch.cuda.Stream owned by the caller) the masks of
    chunk i run beside the up-sampling / chain of chunk i+1 — one is bound by instruction issue, the other by HBM
    writes.  Same kernels, same outputs bit for bit (samples are independent).  `done`: list of >= B/chunk events
    (caller-owned, reused across calls).  On return the current stream has joined `side`."""
    f, b = lo_fwd, lo_bwd
    B, n, _, h, w = f.shape
    H, W = (8 * h, 8 * w) if flow_up else (h, w)
    if out is None:
        out = (torch.empty((B, 2, H, W), device=f.device, dtype=torch.float32),
               torch.empty((B, 2, H, W), device=f.device, dtype=torch.float32),
               torch.empty((B, H, W), device=f.device, dtype=torch.uint8),
               torch.empty((B, H, W), device=f.device, dtype=torch.uint8))
    ff, fb, mf, mb = out
    main = torch.cuda.current_stream(f.device)
    nchunks = (B + chunk - 1) // chunk
    if side is not None and (done is None or len(done) < nchunks):
        done = [torch.cuda.Event() for _ in range(nchunks)]
    for i in range(nchunks):
        s = slice(i * chunk, min(B, (i + 1) * chunk))
        if side is None:
            ops.flow_stage(f[s], b[s], flow_up, alpha_1, alpha_2, False, out=(ff[s], fb[s], mf[s], mb[s]))
            continue
        ops.flow_stage(f[s], b[s], flow_up, None, None, False, out=(ff[s], fb[s], None, None))
        done[i].record(main)
        side.wait_event(done[i])
        with torch.cuda.stream(side):
            ops.fb_masks(ff[s], fb[s], alpha_1, alpha_2, False, out=(mf[s], mb[s]))
    if side is not None:
        main.wait_stream(side)
    return ff, fb, mf.view(torch.bool), mb.view(torch.bool)



B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
chunks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4, 6, 8, 16]
REPS = 12
f, b = synth.flow_fields(B, n, seed=1)
f, b = f.cuda(), b.cuda()
H, W = 8 * f.shape[3], 8 * f.shape[4]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = (torch.empty((B, 2, H, W), device="cuda"), torch.empty((B, 2, H, W), device="cuda"),
       torch.empty((B, H, W), device="cuda", dtype=torch.uint8), torch.empty((B, H, W), device="cuda", dtype=torch.uint8))
side = torch.cuda.Stream(priority=0)
side_hi = torch.cuda.Stream(priority=-1)
events = [torch.cuda.Event() for _ in range(B)]


def sha():
    torch.cuda.synchronize()
    return hashlib.sha256(out[2].cpu().numpy().tobytes() + out[3].cpu().numpy().tobytes()
                          + out[0][::7].cpu().numpy().tobytes()).hexdigest()[:16]


def timed(name, fn_eager):
    for t in out:
        t.zero_()
    for _ in range(3):
        fn_eager()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()  # replayed graph: the comparison is between schedules, not host launch rates
    with torch.cuda.graph(g):
        fn_eager()
    fn = g.replay
    for t in out:
        t.zero_()
    fn()
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:<12} median {ts[len(ts) // 2] * 1000:7.1f} us  min {ts[0] * 1000:7.1f} us  sha {sha()}", flush=True)


timed("whole", lambda: ops.flow_stage(f, b, out=out))
for S in chunks:
    timed(f"chunk{S}", lambda: flow_stage_chunked(f, b, S, out=out))
for S in chunks:
    timed(f"over{S}", lambda: flow_stage_chunked(f, b, S, side=side, done=events, out=out))
for S in chunks:
    timed(f"overhi{S}", lambda: flow_stage_chunked(f, b, S, side=side_hi, done=events, out=out))
