"""Flow store wire format — what feeds the flow stage from disk.

The reference keeps one `.pth` per video and direction: a pickled fp32 tensor `[num_flow, 2, 90, 160]`
(~46 MB for 400 frames) that `load_flow` (contrast/data/dataset.py:341-353) `torch.load`s IN FULL for every
sample and direction just to slice `n_frames-1` links out of it.  `.flw` is the same tensor behind a
64-byte header, so a sample costs one `pread`-sized page-cache copy of the slice it needs:

    offset 0   8s   magic  b"PPFLOW01"
           8   u32  dtype  0 = float32 (bit-exact), 1 = float16 (lossy, opt-in)
          12   u32  ndim   = 4
          16   4*u64 shape  num_flow, 2, h, w
          48   16 bytes reserved (zero)
          64   raw little-endian C-order payload

`load_flow / calc_bwd_idx / load_flows` keep the reference's names, argument meaning, return values and
errors (`NotImplementedError` for an unknown extension); `.pth` files still load the reference's way.
`PinnedFlowStager` turns a list of per-sample slices into the `[B, n, 2, h, w]` pinned batch the flow
stage copies from.  Host-side only; nothing here touches the GPU.
"""
import os
import struct

import numpy as np
import torch

MAGIC = b"PPFLOW01"
HEADER_BYTES = 64
_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("<f2")}


def write_flw(path, flow, dtype="float32"):
    """Write a `[num_flow, 2, h, w]` tensor / array as `.flw`.  dtype "float32" (default, bit-exact) or "float16"."""
    a = flow.detach().cpu().numpy() if isinstance(flow, torch.Tensor) else np.asarray(flow)
    if a.ndim != 4 or a.shape[1] != 2:
        raise ValueError(f"write_flw: expected [num_flow, 2, h, w], got {a.shape}")
    code = {"float32": 0, "float16": 1}[dtype]
    a = np.ascontiguousarray(a, dtype=_DTYPES[code])
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC + struct.pack("<II4Q", code, 4, *a.shape) + b"\0" * 16)
        f.write(a.tobytes())
    os.replace(tmp, path)


def convert_pth_to_flw(src, dst=None, dtype="float32"):
    """One-off conversion of a reference flow file (dataset_prepare/raft_bdd100k output)."""
    dst = dst or os.path.splitext(src)[0] + ".flw"
    write_flw(dst, torch.load(src, map_location="cpu"), dtype)
    return dst


def _open_flw(path):
    with open(path, "rb") as f:
        head = f.read(HEADER_BYTES)
    if len(head) != HEADER_BYTES or head[:8] != MAGIC:
        raise ValueError(f"{path}: not a .flw flow store (bad magic)")
    code, ndim, n, c, h, w = struct.unpack("<II4Q", head[8:48])
    if ndim != 4 or c != 2 or code not in _DTYPES:
        raise ValueError(f"{path}: unsupported .flw header (dtype {code}, ndim {ndim}, shape {(n, c, h, w)})")
    expect = HEADER_BYTES + n * c * h * w * _DTYPES[code].itemsize
    if os.path.getsize(path) != expect:
        raise ValueError(f"{path}: truncated flow store ({os.path.getsize(path)} bytes, header says {expect})")
    return np.memmap(path, dtype=_DTYPES[code], mode="r", offset=HEADER_BYTES, shape=(n, c, h, w))


def load_flow(path, s_idx, n_idx, return_num=True):
    """contrast/data/dataset.py:341-353 with the `.flw` branch added.  Returns flow[s_idx:n_idx] as a
    float32 tensor (python slice semantics, like the reference) and, optionally, num_flow."""
    ext = os.path.splitext(os.path.basename(path))[-1]
    if ext == ".pth":
        flow_tmp = torch.load(path, map_location="cpu")
        num_flow = flow_tmp.shape[0]
        flow = flow_tmp[s_idx:n_idx]
    elif ext == ".flw":
        mm = _open_flw(path)
        num_flow = mm.shape[0]
        flow = torch.from_numpy(np.array(mm[s_idx:n_idx], dtype=np.float32))  # copies only the slice
    else:
        raise NotImplementedError(f"{ext} is not supported!!")
    if return_num:
        return flow, num_flow
    return flow


def calc_bwd_idx(fwd_s_idx, fwd_n_idx, num_flow):
    """contrast/data/dataset.py:356-360: the backward file stores the links in reverse frame order."""
    flow_frames = fwd_n_idx - fwd_s_idx
    bwd_n_idx = num_flow - fwd_s_idx
    return bwd_n_idx - flow_frames, bwd_n_idx


def load_flows(fwd_pathes, bwd_pathes):
    """contrast/data/dataset.py:363-369: (path, s_idx, n_idx) triples -> (flow_fwd, flow_bwd)."""
    _, fwd_s_idx, fwd_n_idx = fwd_pathes
    bwd_path = bwd_pathes[0]
    flow_fwd, num_flow = load_flow(*fwd_pathes, return_num=True)
    bwd_s_idx, bwd_n_idx = calc_bwd_idx(fwd_s_idx, fwd_n_idx, num_flow)
    return flow_fwd, load_flow(bwd_path, bwd_s_idx, bwd_n_idx, return_num=False)


class PinnedFlowStager:
    """Collates per-sample link slices into reusable pinned `[B, n, 2, h, w]` batches (double-buffered), the
    layout `pixpro_b200.ops.flow_stage` / `HostPixelStep` copy from with non_blocking=True."""

    def __init__(self, batch, n_links, h=90, w=160, buffers=2):
        pin = torch.cuda.is_available()
        mk = lambda: torch.empty((batch, n_links, 2, h, w), dtype=torch.float32, pin_memory=pin)
        self.bufs = [(mk(), mk()) for _ in range(buffers)]
        self.i = 0

    def collate(self, samples):
        """samples: list of (flow_fwd, flow_bwd) per sample, each [n, 2, h, w].  Returns the pinned pair."""
        fwd, bwd = self.bufs[self.i % len(self.bufs)]
        self.i += 1
        if len(samples) != fwd.shape[0]:
            raise ValueError(f"PinnedFlowStager: expected {fwd.shape[0]} samples, got {len(samples)}")
        for b, (f, g) in enumerate(samples):
            fwd[b].copy_(f)
            bwd[b].copy_(g)
        return fwd, bwd
