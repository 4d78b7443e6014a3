"""Multi-tensor optimizer-side entries of libpixpro_b200.so:

  ema_update(pairs, m)   EMA of the key branch, contrast/models/PixPro.py:322-337 — one launch
  lars_sgd_step(...)     LARS.step() around torch.optim.SGD, contrast/lars.py:109-152 — three launches

A parameter set is handed to the library as a device table of PpMtTensor records plus a chunk
map (include/pixpro_b200.h).  The records hold raw pointers, so the table is rebuilt whenever a
pointer or a hyper-parameter changes (gradients are reallocated by zero_grad(set_to_none=True))
and cached otherwise.  Nothing here computes: all arithmetic is in the kernels; CPU tensors raise.
"""
import numpy as np
import torch

from . import _cabi

MT_LARS = 1
MT_FIRST_STEP = 2

_REC = np.dtype([("a", "<u8"), ("b", "<u8"), ("c", "<u8"), ("numel", "<i8"), ("s0", "<f4"), ("s1", "<f4"), ("s2", "<f4"),
                 ("s3", "<f4"), ("flags", "<i4"), ("pad", "<i4")])
assert _REC.itemsize == 56


def _check(t, what):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise _cabi.PixProB200Error(f"{what}: expected a contiguous CUDA float32 tensor (no CPU fallback), got "
                                    f"{t.device} {t.dtype} contiguous={t.is_contiguous()}")


class _TensorSet:
    """Device-side table + chunk map of one list of tensor records, cached on its content."""

    def __init__(self):
        self.key = None
        self.table = self.cmap = self.first = None
        self.nchunks = 0

    def update(self, recs, device):
        key = recs.tobytes()
        if key == self.key:
            return
        chunk = _cabi.lib().pp_mt_chunk_elems()
        per = (recs["numel"] + chunk - 1) // chunk
        first = np.zeros(len(recs) + 1, np.int32)
        np.cumsum(per, out=first[1:])
        self.nchunks = int(first[-1])
        cmap = np.empty((self.nchunks, 2), np.int32)
        cmap[:, 0] = np.repeat(np.arange(len(recs), dtype=np.int32), per)
        cmap[:, 1] = np.arange(self.nchunks, dtype=np.int32) - np.repeat(first[:-1], per)
        self.table = torch.from_numpy(recs.view(np.uint8).copy()).to(device)
        self.cmap = torch.from_numpy(cmap).to(device)
        self.first = torch.from_numpy(first).to(device)
        self.key = key


_ema_sets = {}


def ema_update(pairs, momentum, cache_key=None):
    """k <- k*m + q*(1-m) for every (q, k) pair (online parameter, momentum parameter), in place, one launch.
    `momentum` is the python float the reference computes; 1-m is formed in double, as it does."""
    pairs = list(pairs)
    if not pairs:
        return
    recs = np.zeros(len(pairs), _REC)
    for i, (q, k) in enumerate(pairs):
        _check(q, "ema_update q")
        _check(k, "ema_update k")
        if q.shape != k.shape:
            raise ValueError("ema_update: shape mismatch")
        recs[i]["a"], recs[i]["b"], recs[i]["numel"] = q.data_ptr(), k.data_ptr(), q.numel()
    dev = pairs[0][1].device
    ts = _ema_sets.setdefault((cache_key, dev), _TensorSet())
    ts.update(recs, dev)
    m = float(momentum)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().pp_ema_update(ts.table.data_ptr(), ts.cmap.data_ptr(), ts.nchunks, m, 1. - m,
                                              torch.cuda.current_stream(dev).cuda_stream), "pp_ema_update")


class LarsSgdStep:
    """The fused LARS + SGD step over a fixed optimizer (see contrast/lars.py of this package)."""

    def __init__(self):
        self.ts = _TensorSet()
        self.ws = None

    def __call__(self, entries, trust_coef, eps):
        """entries: list of (param, grad, momentum_buffer or None, weight_decay, lr, momentum, dampening, lars, first)."""
        if not entries:
            return
        recs = np.zeros(len(entries), _REC)
        for i, (p, g, buf, wd, lr, mom, damp, lars, first) in enumerate(entries):
            _check(p, "lars_sgd_step param")
            _check(g, "lars_sgd_step grad")
            if buf is not None:
                _check(buf, "lars_sgd_step momentum buffer")
            r = recs[i]
            r["a"], r["b"], r["c"], r["numel"] = p.data_ptr(), g.data_ptr(), (buf.data_ptr() if buf is not None else 0), p.numel()
            r["s0"], r["s1"], r["s2"], r["s3"] = wd, lr, mom, damp
            r["flags"] = (MT_LARS if lars else 0) | (MT_FIRST_STEP if first else 0)
        dev = entries[0][0].device
        self.ts.update(recs, dev)
        L = _cabi.lib()
        need = L.pp_lars_workspace(len(entries), self.ts.nchunks)
        if self.ws is None or self.ws.numel() < need or self.ws.device != dev:
            self.ws = torch.empty((need,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(L.pp_lars_sgd_step(self.ts.table.data_ptr(), len(entries), self.ts.cmap.data_ptr(), self.ts.first.data_ptr(),
                                           self.ts.nchunks, float(trust_coef), float(eps), self.ws.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream), "pp_lars_sgd_step")
