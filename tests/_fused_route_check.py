"""Helper of test_gpu_parity.py::test_flow_stage_fused_routes: run in a fresh process with PIXPRO_B200_FBUP set (the library
reads the switch once), checks the n = 1 flow stage of that route against the reference's goldens (720x1280: the compile-time
frame-size instance) and against the CPU oracle on small frames (generic instance, flows crossing the border)."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, unpack_mask  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from pixpro_b200 import _cabi, ops, synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


g = load_golden("flow_stage_full_n1")
_cabi.profile_enable(True)
ff, fb, mf, mb = ops.flow_stage(torch.from_numpy(g["lo_fwd"]).cuda(), torch.from_numpy(g["lo_bwd"]).cuda())
torch.cuda.synchronize()
names = set(_cabi.profile_report())
_cabi.profile_enable(False)
mode = os.environ.get("PIXPRO_B200_FBUP")
want = {"0": {"chain_up", "fb"}, "1": {"chain_up1", "fb_up_w", "fb_up"}, "2": {"chain_up1", "fb_up_w", "fb1"}}[mode]
assert names == want, (mode, names)
assert sha(ff.cpu().numpy()) == str(g["flow_fwd_sha"]) and sha(fb.cpu().numpy()) == str(g["flow_bwd_sha"])
assert np.array_equal(mf.cpu().numpy(), unpack_mask(g["mask_fwd"], mf.shape))
assert np.array_equal(mb.cpu().numpy(), unpack_mask(g["mask_bwd"], mb.shape))
import random
rng = random.Random(2026)
fuzz = [(rng.randint(1, 3), 6 * rng.randint(1, 5), 8 * rng.randint(1, 4), rng.choice([0.5, 3.0, 10.0, 40.0]), 100 + i) for i in range(8)]
for B, h, w, mag, seed in [(3, 12, 16, 6.0, 5), (2, 18, 24, 1.0, 6), (2, 24, 16, 30.0, 7), (5, 90, 160, 1.5, 8)] + fuzz:
    f, b = synth.flow_fields(B, 1, h=h, w=w, seed=seed, magnitude=mag)
    ref = orc.flow_stage(f.numpy(), b.numpy(), flow_up=True)
    got = ops.flow_stage(f.cuda(), b.cuda())
    for name, gt, wt in zip(["flow_fwd", "flow_bwd", "mask_fwd", "mask_bwd"], got, ref):
        a = gt.cpu().numpy()
        assert a.shape == wt.shape and np.array_equal(a.view(np.uint8), np.ascontiguousarray(wt).view(np.uint8)), (name, B, h, w)
_cabi.fb_redo_count()  # raises if an mbarrier wait timed out
print("FUSED ROUTE OK", mode)
