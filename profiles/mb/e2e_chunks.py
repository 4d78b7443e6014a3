"""End-to-end (pinned host buffers in / out) step time of HostPixelStep against the chunking of the link copies.
Round 1 (equal chunks, n_frames=2, B=64): 1 chunk 1.138 ms, 2: 1.021, 4: 0.984, 8: 1.099, 16: 1.046; sparse: 0.839.
Round 2 adds chunk-size lists (a small first chunk starts the flow kernels early, large later chunks run in the kernels' efficient regime)."""
import os, sys, types, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
import bench
from pixpro_b200.host_step import HostPixelStep
a = types.SimpleNamespace(batch=64, n_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 2, grid=7)
dev = torch.device("cuda:0")
host = bench.make_inputs(a.batch, a.n_frames, a.grid, 1234)
pinned = {k: v.pin_memory() for k, v in host.items()}
w, bias = host["w"].to(dev), host["bias"].to(dev)
keys = ["feat1", "feat2", "k1", "k2", "c1", "c2", "lo_f", "lo_b"]
hin = {k: pinned[k] for k in keys}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SWEEP = [(4, False), ("auto", False), ([8, 56], False), ([16, 48], False), ([4, 12, 48], False), ([8, 16, 40], False), ([4, 28, 32], False), (2, False), (1, True)]
if len(sys.argv) > 2:  # second sweep
    SWEEP = [("auto", False), ([8, 8, 16, 32], False), ([4, 12, 16, 32], False), ([6, 18, 40], False), ([10, 22, 32], False), ([8, 20, 36], False), ([12, 20, 32], False), ([8, 24, 32], False)]
for chunks, sparse in SWEEP:
    st = HostPixelStep(dev, a.batch, 256, a.grid, flow_chunks=chunks, sparse=sparse)
    for _ in range(5): st(hin, w, bias)
    ts = []
    for _ in range(20):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); st(hin, w, bias); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"n_frames={a.n_frames} flow_chunks={chunks} sparse={sparse}: {sum(ts)/len(ts):.3f} ms/step (min {min(ts):.3f})")
