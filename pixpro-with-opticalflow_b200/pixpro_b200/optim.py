"""Multi-tensor optimizer-side entries of libpixpro_b200.so:

  ema_update(pairs, m)   EMA of the key branch, contrast/models/PixPro.py:322-337 — one launch
  lars_sgd_step(...)     LARS.step() around torch.optim.SGD, contrast/lars.py:109-152 — three launches

A parameter set is handed to the library as a device table of PpMtTensor records plus a chunk
map (include/pixpro_b200.h).  The records hold raw pointers: table and chunk map are rebuilt only when a
pointer moves (gradients reallocated by zero_grad(set_to_none=True)); a change of the per-step scalars alone
(the scheduler moves the learning rate every iteration) rewrites 56 bytes per tensor through a pinned staging
ring, asynchronously, without re-validating or synchronising (see _TensorSet).  Nothing here computes: all
arithmetic is in the kernels; CPU tensors raise.
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


def _all_ok(tensors, what):
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            _check(t, what)


class _TensorSet:
    """Device-side table + chunk map of one list of tensor records.

    Two host-side signatures decide what a step costs (the learning rate changes every iteration):
      * the POINTER signature (raw pointers, element counts): when it changes the tensors are re-validated and
        table + chunk map are rebuilt and uploaded;
      * the SCALAR signature (weight decay / lr / momentum / dampening / flags per record): when only it changes, the
        scalar columns of the host copy of the table are rewritten and the table alone (56 B per tensor) is copied
        over the SAME device memory — from a small ring of pinned staging buffers with non_blocking=True, ordered on
        the current stream ahead of the kernels that read it.  No validation pass, no synchronisation, no reallocation."""
    RING = 4

    def __init__(self):
        self.ptr_sig = self.scal_sig = None
        self.table = self.cmap = self.first = None
        self.nchunks = 0
        self.n = 0
        self._stage, self._events, self._slot = [], [], 0
        self.rebuilds = self.scalar_updates = 0   # diagnostics (tests / benches)

    def _upload_table(self, recs, device):
        """recs -> self.table through a pinned staging slot (asynchronous on the current stream)."""
        nbytes = recs.nbytes
        if not self._stage or self._stage[0].numel() != nbytes:
            pin = torch.cuda.is_available()
            self._stage = [torch.empty((nbytes,), dtype=torch.uint8, pin_memory=pin) for _ in range(self.RING)]
            self._events = [None] * self.RING
        i = self._slot
        self._slot = (i + 1) % self.RING
        if self._events[i] is not None:
            self._events[i].synchronize()   # the copy that last used this slot (RING steps ago) has long finished
        self._stage[i].numpy()[:] = recs.view(np.uint8).reshape(-1)
        self.table.copy_(self._stage[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self._events[i] = ev

    def update(self, ptr_sig, scal_sig, build, fill_scalars, device):
        if ptr_sig != self.ptr_sig:
            recs = build()
            fill_scalars(recs)
            chunk = _cabi.lib().pp_mt_chunk_elems()
            per = (recs["numel"] + chunk - 1) // chunk
            first = np.zeros(len(recs) + 1, np.int32)
            np.cumsum(per, out=first[1:])
            self.nchunks = int(first[-1])
            self.n = len(recs)
            cmap = np.empty((self.nchunks, 2), np.int32)
            cmap[:, 0] = np.repeat(np.arange(len(recs), dtype=np.int32), per)
            cmap[:, 1] = np.arange(self.nchunks, dtype=np.int32) - np.repeat(first[:-1], per)
            with torch.cuda.device(device):
                self.table = torch.empty((recs.nbytes,), dtype=torch.uint8, device=device)
                rest = np.concatenate([cmap.view(np.uint8).ravel(), first.view(np.uint8)])
                dev_rest = torch.from_numpy(rest).to(device)   # rare path (pointers moved): a plain blocking upload
                self.cmap, self.first = dev_rest[:cmap.nbytes], dev_rest[cmap.nbytes:]
                self._stage = []
                self._upload_table(recs, device)
            self.recs = recs
            self.ptr_sig, self.scal_sig = ptr_sig, scal_sig
            self.rebuilds += 1
        elif scal_sig != self.scal_sig:
            fill_scalars(self.recs)
            with torch.cuda.device(device):
                self._upload_table(self.recs, device)
            self.scal_sig = scal_sig
            self.scalar_updates += 1


_ema_sets = {}


def ema_update(pairs, momentum, cache_key=None):
    """k <- k*m + q*(1-m) for every (q, k) pair (online parameter, momentum parameter), in place, one launch.
    `momentum` is the python float the reference computes; 1-m is formed in double, as it does."""
    pairs = list(pairs)
    if not pairs:
        return
    qs, ks = [q for q, _ in pairs], [k for _, k in pairs]
    sig = (tuple(t.data_ptr() for t in qs), tuple(t.data_ptr() for t in ks), tuple(t.numel() for t in ks))
    dev = ks[0].device
    if not ks[0].is_cuda:
        _check(ks[0], "ema_update k")
    ts = _ema_sets.setdefault((cache_key, dev), _TensorSet())

    def build():
        _all_ok(qs, "ema_update q")
        _all_ok(ks, "ema_update k")
        if any(q.shape != k.shape for q, k in pairs):
            raise ValueError("ema_update: shape mismatch")
        recs = np.zeros(len(pairs), _REC)
        recs["a"], recs["b"], recs["numel"] = sig[0], sig[1], sig[2]
        return recs

    ts.update(sig, None, build, lambda recs: None, dev)
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
        ps, gs, bs = [e[0] for e in entries], [e[1] for e in entries], [e[2] for e in entries]
        ptr_sig = (tuple(t.data_ptr() for t in ps), tuple(t.data_ptr() for t in gs), tuple(0 if b is None else b.data_ptr() for b in bs),
                   tuple(t.numel() for t in ps))
        scal_sig = tuple(e[3:] for e in entries)
        dev = ps[0].device
        if not ps[0].is_cuda:
            _check(ps[0], "lars_sgd_step param")

        def build():
            _all_ok(ps, "lars_sgd_step param")
            _all_ok(gs, "lars_sgd_step grad")
            _all_ok([b for b in bs if b is not None], "lars_sgd_step momentum buffer")
            recs = np.zeros(len(entries), _REC)
            recs["a"], recs["b"], recs["c"] = ptr_sig[0], ptr_sig[1], ptr_sig[2]
            recs["numel"] = ptr_sig[3]
            return recs

        def fill_scalars(recs):
            recs["s0"], recs["s1"] = [e[3] for e in entries], [e[4] for e in entries]
            recs["s2"], recs["s3"] = [e[5] for e in entries], [e[6] for e in entries]
            recs["flags"] = [(MT_LARS if e[7] else 0) | (MT_FIRST_STEP if e[8] else 0) for e in entries]

        self.ts.update(ptr_sig, scal_sig, build, fill_scalars, dev)
        L = _cabi.lib()
        need = L.pp_lars_workspace(len(entries), self.ts.nchunks)
        if self.ws is None or self.ws.numel() < need or self.ws.device != dev:
            self.ws = torch.empty((need,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(L.pp_lars_sgd_step(self.ts.table.data_ptr(), len(entries), self.ts.cmap.data_ptr(), self.ts.first.data_ptr(),
                                           self.ts.nchunks, float(trust_coef), float(eps), self.ws.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream), "pp_lars_sgd_step")
