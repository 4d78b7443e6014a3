"""Flow store (.flw) — SURVEY §8(f) rank 2.  CPU only: the format, the drop-in load_flow/load_flows
(contrast/data/dataset.py:341-369) and the pinned collation."""
import os
import sys
import time

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from pixpro_b200 import flowstore as fs  # noqa: E402


@pytest.fixture()
def video(tmp_path):
    g = torch.Generator().manual_seed(4)
    fwd = torch.randn(37, 2, 18, 32, generator=g)
    bwd = torch.randn(37, 2, 18, 32, generator=g)
    p = {}
    for name, t in (("fwd", fwd), ("bwd", bwd)):
        p[name + ".pth"] = str(tmp_path / f"{name}.pth")
        torch.save(t, p[name + ".pth"])
        p[name + ".flw"] = fs.convert_pth_to_flw(p[name + ".pth"])
    return fwd, bwd, p


def reference_load_flows(fwd, bwd, s, n):
    """What contrast/data/dataset.py:341-369 returns for in-memory tensors."""
    num = fwd.shape[0]
    frames = n - s
    bn = num - s
    return fwd[s:n], bwd[bn - frames:bn]


@pytest.mark.parametrize("s,n", [(0, 5), (3, 8), (32, 37), (36, 37), (35, 40), (0, 0)])
def test_flw_slices_equal_the_pth_path_bit_for_bit(video, s, n):
    fwd, bwd, p = video
    a, num_a = fs.load_flow(p["fwd.pth"], s, n)
    b, num_b = fs.load_flow(p["fwd.flw"], s, n)
    assert num_a == num_b == 37 and a.dtype == b.dtype == torch.float32
    assert torch.equal(a, b) and a.shape == b.shape
    ff, fb = fs.load_flows((p["fwd.flw"], s, n), (p["bwd.flw"], s, n))
    rf, rb = reference_load_flows(fwd, bwd, s, n)
    assert torch.equal(ff, rf) and torch.equal(fb, rb)


def test_unknown_extension_and_corrupt_files(video, tmp_path):
    _, _, p = video
    with pytest.raises(NotImplementedError):
        fs.load_flow(str(tmp_path / "x.npy"), 0, 1)
    bad = str(tmp_path / "bad.flw")
    open(bad, "wb").write(b"NOTAFLOW" + b"\0" * 56)
    with pytest.raises(ValueError):
        fs.load_flow(bad, 0, 1)
    trunc = str(tmp_path / "trunc.flw")
    open(trunc, "wb").write(open(p["fwd.flw"], "rb").read()[:-8])
    with pytest.raises(ValueError):
        fs.load_flow(trunc, 0, 1)
    with pytest.raises(ValueError):
        fs.write_flw(str(tmp_path / "y.flw"), torch.zeros(3, 3, 4, 4))


def test_float16_store_is_opt_in_and_close(video, tmp_path):
    fwd, _, _ = video
    q = str(tmp_path / "h.flw")
    fs.write_flw(q, fwd, dtype="float16")
    x = fs.load_flow(q, 2, 7, return_num=False)
    assert x.dtype == torch.float32 and torch.equal(x, fwd[2:7].half().float())
    assert os.path.getsize(q) == fs.HEADER_BYTES + fwd.numel() * 2


def test_pinned_stager_collates_the_loader_layout(video):
    fwd, bwd, p = video
    st = fs.PinnedFlowStager(batch=3, n_links=5, h=18, w=32)
    samples = [fs.load_flows((p["fwd.flw"], s, s + 5), (p["bwd.flw"], s, s + 5)) for s in (0, 4, 30)]
    F, B = st.collate(samples)
    assert F.shape == (3, 5, 2, 18, 32) and B.shape == F.shape
    for i, s in enumerate((0, 4, 30)):
        rf, rb = reference_load_flows(fwd, bwd, s, s + 5)
        assert torch.equal(F[i], rf) and torch.equal(B[i], rb)
    F2, _ = st.collate(samples)
    assert F2.data_ptr() != F.data_ptr()  # double-buffered: the previous batch may still be in flight
    with pytest.raises(ValueError):
        st.collate(samples[:2])


def test_slice_read_is_much_cheaper_than_loading_the_file(tmp_path):
    """A 400-frame video at the published 90x160 (46 MB): reading 5 links from .flw vs torch.load of the .pth."""
    t = torch.randn(400, 2, 90, 160)
    pth = str(tmp_path / "v.pth")
    torch.save(t, pth)
    flw = fs.convert_pth_to_flw(pth)

    def best(fn, reps=5):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    t_pth = best(lambda: fs.load_flow(pth, 100, 105))
    t_flw = best(lambda: fs.load_flow(flw, 100, 105))
    assert torch.equal(fs.load_flow(pth, 100, 105)[0], fs.load_flow(flw, 100, 105)[0])
    print(f"load 5 of 400 links: .pth {t_pth * 1e3:.2f} ms, .flw {t_flw * 1e3:.3f} ms ({t_pth / t_flw:.0f}x)")
    assert t_flw * 5 < t_pth
