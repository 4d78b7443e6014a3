"""torch-facing wrappers of the C ABI (include/pixpro_b200.h).

torch supplies device memory (the caching allocator) and the current CUDA stream; all
arithmetic runs in libpixpro_b200.so.  Every wrapper validates device / dtype / layout and
raises — there is no CPU or eager-PyTorch fallback.
"""
import os

import torch

from . import _cabi

_DIV_MODES = {"ieee": 0, "rcp": 1}
_div_mode = _DIV_MODES[os.environ.get("PIXPRO_B200_DIV", "ieee")]


def set_div_mode(mode):
    """'ieee': tensor/scalar sites round like the reference on CPU (true division; the pinned
    oracle).  'rcp': like torch's CUDA true-divide kernel (x * fl32(1/s))."""
    global _div_mode
    _div_mode = _DIV_MODES[mode]


def get_div_mode():
    return {v: k for k, v in _DIV_MODES.items()}[_div_mode]


_serial = False


def set_serial(on):
    """True: ops issue every launch on the current stream (no internal side streams) — for per-kernel timing
    passes, where an overlapped launch's device time would include its overlap partner's."""
    global _serial
    _serial = bool(on)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise _cabi.PixProB200Error(f"{name}: tensor is on {t.device}; the pixel-pretext kernels are CUDA-only "
                                    "(no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _f32(t, name):
    if isinstance(t, torch.Tensor) and t.is_cuda and t.dtype in (torch.float16, torch.bfloat16, torch.float64):
        t = t.float()
    return _req(t, name)


def _mask_u8(m, name):
    if m.dtype == torch.bool:
        m = m.contiguous().view(torch.uint8)
    return _req(m, name, torch.uint8)


def _ptr(t):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------ flow --

def upflow8(flow):
    """contrast/flow/utils/utils.py:87-89 (bilinear mode)."""
    flow = _f32(flow, "flow")
    assert flow.ndim == 4, "upflow8 expects [N,C,h,w]"
    N, C, h, w = flow.shape
    out = torch.empty((N, C, 8 * h, 8 * w), device=flow.device, dtype=torch.float32)
    with torch.cuda.device(flow.device):
        _cabi.check(_cabi.lib().pp_upflow8(_ptr(flow), N * C, h, w, _ptr(out), _stream()), "pp_upflow8")
    return out


def _normalize(x, kind, name):
    x = _f32(x, name)
    assert x.ndim == 4 and x.shape[1] == 2, f"{name} expects [B,2,H,W]"
    B, _, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().pp_normalize(_ptr(x), B, H, W, kind, _div_mode, _ptr(out), _stream()), "pp_normalize")
    return out


def normalize_coord(coords):
    """contrast/util.py:334-339"""
    return _normalize(coords, 0, "coords")


def normalize_flow(flow):
    """contrast/util.py:343-348"""
    return _normalize(flow, 1, "flow")


def denormalize_flow(flow_norm):
    """contrast/util.py:352-357"""
    return _normalize(flow_norm, 2, "flow_norm")


def concat_flow(flows, is_norm=False):
    """contrast/util.py:301-330.  flows [n,B,2,H,W] (any strides over n and B, e.g. the permuted
    loader tensor) -> [B,2,H,W]."""
    if not (isinstance(flows, torch.Tensor) and flows.is_cuda):
        raise _cabi.PixProB200Error("flows: CUDA tensor required (no CPU fallback)")
    assert flows.ndim == 5 and flows.shape[2] == 2, "concat_flow expects [n,B,2,H,W]"
    if flows.dtype != torch.float32:
        flows = flows.float()
    n, B, _, H, W = flows.shape
    if flows.stride()[2:] != (H * W, W, 1):
        flows = flows.contiguous()
    out = torch.empty((B, 2, H, W), device=flows.device, dtype=torch.float32)
    with torch.cuda.device(flows.device):
        _cabi.check(_cabi.lib().pp_concat_flow(_ptr(flows), n, B, H, W, flows.stride(0), flows.stride(1), int(is_norm),
                                               _div_mode, _ptr(out), _stream()), "pp_concat_flow")
    return out


def forward_backward_consistency(flow_fwd, flow_bwd, alpha_1=0.01, alpha_2=0.5, is_norm=False, want_cycle=True,
                                 want_coords=True):
    """contrast/util.py:253-297 -> (coords1_norm or None, mask bool [B,H,W], cycle or None)."""
    f = _f32(flow_fwd, "flow_fwd")
    b = _f32(flow_bwd, "flow_bwd")
    assert f.shape == b.shape and f.ndim == 4 and f.shape[1] == 2
    B, _, H, W = f.shape
    mask = torch.empty((B, H, W), device=f.device, dtype=torch.uint8)
    cyc = torch.empty_like(f) if want_cycle else None
    c1 = torch.empty_like(f) if want_coords else None
    with torch.cuda.device(f.device):
        _cabi.check(_cabi.lib().pp_fb_consistency(_ptr(f), _ptr(b), B, H, W, float(alpha_1), float(alpha_2), int(is_norm),
                                                  _div_mode, _ptr(mask), _ptr(cyc), _ptr(c1), _stream()),
                    "pp_fb_consistency")
    return c1, mask.view(torch.bool), cyc


def flow_stage(lo_fwd, lo_bwd, flow_up=True, alpha_1=0.01, alpha_2=0.5, is_norm=False, out=None):
    """Flow stage of contrast/util.py:175-248 (use_flow_file, not use_flow_frames), fused.

    lo_fwd/lo_bwd: loader layout [B,n,2,h,w].  Returns (flow_fwd, flow_bwd [B,2,H,W],
    mask_fwd, mask_bwd bool [B,H,W] or None when alpha_1/alpha_2 is None).
    out=(flow_fwd, flow_bwd, mask_fwd u8, mask_bwd u8): write into preallocated contiguous tensors
    (e.g. batch slices of larger buffers) instead of allocating."""
    f = _f32(lo_fwd, "lo_fwd")
    b = _f32(lo_bwd, "lo_bwd")
    assert f.ndim == 5 and f.shape == b.shape and f.shape[2] == 2, "flow_stage expects [B,n,2,h,w]"
    B, n, _, h, w = f.shape
    H, W = (8 * h, 8 * w) if flow_up else (h, w)
    use_mask = alpha_1 is not None and alpha_2 is not None
    if out is not None:
        ff, fb, mf, mb = out
        for t_, shp, dt in ((ff, (B, 2, H, W), torch.float32), (fb, (B, 2, H, W), torch.float32)) + \
                (((mf, (B, H, W), torch.uint8), (mb, (B, H, W), torch.uint8)) if use_mask else ()):
            if tuple(t_.shape) != shp or t_.dtype != dt or not t_.is_contiguous() or not t_.is_cuda:
                raise ValueError(f"flow_stage: out tensor must be contiguous CUDA {dt} of shape {shp}")
    else:
        ff = torch.empty((B, 2, H, W), device=f.device, dtype=torch.float32)
        fb = torch.empty((B, 2, H, W), device=f.device, dtype=torch.float32)
        mf = torch.empty((B, H, W), device=f.device, dtype=torch.uint8) if use_mask else None
        mb = torch.empty((B, H, W), device=f.device, dtype=torch.uint8) if use_mask else None
    L = _cabi.lib()
    wsz = L.pp_flow_stage_workspace(B, n, h, w, int(flow_up))  # 0: the fused chain kernel needs no scratch
    ws = torch.empty((wsz,), device=f.device, dtype=torch.uint8) if wsz else None
    with torch.cuda.device(f.device):
        _cabi.check(L.pp_flow_stage(_ptr(f), _ptr(b), B, n, h, w, int(flow_up), int(use_mask),
                                    float(alpha_1 or 0.0), float(alpha_2 or 0.0), int(is_norm), _div_mode,
                                    _ptr(ff), _ptr(fb), _ptr(mf), _ptr(mb), _ptr(ws), wsz, _stream()), "pp_flow_stage")
    return ff, fb, (mf.view(torch.bool) if use_mask else None), (mb.view(torch.bool) if use_mask else None)


def fb_masks(flow_fwd, flow_bwd, alpha_1=0.01, alpha_2=0.5, is_norm=False, out=None):
    """Both FB masks of apply_optical_flow (contrast/util.py:211-213) in one launch (pp_fb_masks).
    flow_fwd/flow_bwd [B,2,H,W] -> (mask_fwd, mask_bwd) u8 [B,H,W] (out=(mf, mb) to write in place)."""
    f = _f32(flow_fwd, "flow_fwd")
    b = _f32(flow_bwd, "flow_bwd")
    assert f.shape == b.shape and f.ndim == 4 and f.shape[1] == 2 and f.is_contiguous() and b.is_contiguous()
    B, _, H, W = f.shape
    if out is not None:
        mf, mb = out
        for t_ in (mf, mb):
            if tuple(t_.shape) != (B, H, W) or t_.dtype != torch.uint8 or not t_.is_contiguous() or not t_.is_cuda:
                raise ValueError(f"fb_masks: out tensor must be contiguous CUDA uint8 of shape {(B, H, W)}")
    else:
        mf = torch.empty((B, H, W), device=f.device, dtype=torch.uint8)
        mb = torch.empty((B, H, W), device=f.device, dtype=torch.uint8)
    with torch.cuda.device(f.device):
        _cabi.check(_cabi.lib().pp_fb_masks(_ptr(f), _ptr(b), B, H, W, float(alpha_1), float(alpha_2), int(is_norm),
                                            _div_mode, _ptr(mf), _ptr(mb), _stream()), "pp_fb_masks")
    return mf, mb


def calc_mask_ratio(mask):
    """contrast/util.py:361-366"""
    m = _mask_u8(mask, "mask")
    assert m.ndim == 3
    B, H, W = m.shape
    out = torch.empty((B,), device=m.device, dtype=torch.float32)
    with torch.cuda.device(m.device):
        _cabi.check(_cabi.lib().pp_calc_mask_ratio(_ptr(m), B, H, W, _ptr(out), _stream()), "pp_calc_mask_ratio")
    return out


# --------------------------------------------------------------------- sparse correspondence --

def sparse_corr(lo_fwd, lo_bwd, coord_fwd, coord_bwd, grid, size, flow_up=True, alpha_1=0.01, alpha_2=0.5):
    """Flow stage + add_optical_flow evaluated only at the G*G grid centres of each sample
    (pp_sparse_corr): bit-identical to flow_stage() followed by add_optical_flow(), without the
    dense [B,2,H,W] composites and [B,H,W] masks.

    lo_fwd/lo_bwd [B,n,2,h,w]; coord_fwd / coord_bwd [B,10] (either may be None): crop descriptors
    of the view warped by the forward / backward composite; size = (H_orig, W_orig).
    Returns (warped_fwd, warped_bwd), each [3,B,P] = (x, y, mask bit as 0/1) or None."""
    f = _f32(lo_fwd, "lo_fwd")
    b = _f32(lo_bwd, "lo_bwd")
    assert f.ndim == 5 and f.shape == b.shape and f.shape[2] == 2, "sparse_corr expects [B,n,2,h,w]"
    B, n, _, h, w = f.shape
    H_orig, W_orig = _size_hw(size)
    use_mask = alpha_1 is not None and alpha_2 is not None
    P = grid * grid
    cf = _f32(coord_fwd, "coord_fwd") if coord_fwd is not None else None
    cb = _f32(coord_bwd, "coord_bwd") if coord_bwd is not None else None
    for c in (cf, cb):
        assert c is None or c.shape == (B, 10)
    wf = torch.empty((3, B, P), device=f.device, dtype=torch.float32) if cf is not None else None
    wb = torch.empty((3, B, P), device=f.device, dtype=torch.float32) if cb is not None else None
    with torch.cuda.device(f.device):
        _cabi.check(_cabi.lib().pp_sparse_corr(_ptr(f), _ptr(b), B, n, h, w, int(flow_up), int(use_mask), float(alpha_1 or 0.0),
                                               float(alpha_2 or 0.0), _ptr(cf), _ptr(cb), grid, H_orig, W_orig, _div_mode,
                                               _ptr(wf), _ptr(wb), _stream()), "pp_sparse_corr")
    return wf, wb


class LazyFlowPair:
    """The flow stage of apply_optical_flow, deferred.  Holds the loader's low-res links; the loss takes
    the sparse correspondence path (sparse_corr) from them, and anything that needs the dense tensors
    (`.dense()`, mask-ratio logging) materialises them once through flow_stage() — same bits either way."""

    def __init__(self, lo_fwd, lo_bwd, flow_up=True, alpha_1=0.01, alpha_2=0.5):
        self.lo_fwd, self.lo_bwd = _f32(lo_fwd, "lo_fwd"), _f32(lo_bwd, "lo_bwd")
        self.flow_up, self.alpha_1, self.alpha_2 = flow_up, alpha_1, alpha_2
        self._dense = None
        B, n, _, h, w = self.lo_fwd.shape
        self.flow_shape = (B, 2, 8 * h, 8 * w) if flow_up else (B, 2, h, w)
        self.use_mask = alpha_1 is not None and alpha_2 is not None
        self.flow = (LazyFlow(self, 0), LazyFlow(self, 1))
        self.mask = (LazyMask(self, 0), LazyMask(self, 1)) if self.use_mask else (None, None)

    def dense(self):
        if self._dense is None:
            self._dense = flow_stage(self.lo_fwd, self.lo_bwd, flow_up=self.flow_up, alpha_1=self.alpha_1, alpha_2=self.alpha_2)
        return self._dense


class LazyFlow:
    """Stands in for one composite flow tensor [B,2,H,W] of apply_optical_flow's return value."""

    def __init__(self, pair, direction):
        self.pair, self.direction = pair, direction

    @property
    def shape(self):
        return torch.Size(self.pair.flow_shape)

    def dense(self):
        return self.pair.dense()[self.direction]

    def clone(self):
        return self


class LazyMask:
    """Stands in for one FB mask [B,H,W] of apply_optical_flow's return value."""

    def __init__(self, pair, direction):
        self.pair, self.direction = pair, direction

    @property
    def shape(self):
        B, _, H, W = self.pair.flow_shape
        return torch.Size((B, H, W))

    def dense(self):
        return self.pair.dense()[2 + self.direction]

    def clone(self):
        return self


def _lazy_warped(flow, mask, coord_q, grid, size):
    """warped centres [3,B,P] of one loss direction from a LazyFlow (and its LazyMask or None)."""
    pr = flow.pair
    if mask is not None and not (isinstance(mask, LazyMask) and mask.pair is pr and mask.direction == flow.direction):
        raise ValueError("a LazyFlow must come with the LazyMask of the same apply_optical_flow call (or None)")
    a1, a2 = (pr.alpha_1, pr.alpha_2) if mask is not None else (None, None)
    cf, cb = (coord_q, None) if flow.direction == 0 else (None, coord_q)
    wf, wb = sparse_corr(pr.lo_fwd, pr.lo_bwd, cf, cb, grid, size, flow_up=pr.flow_up, alpha_1=a1, alpha_2=a2)
    return wf if flow.direction == 0 else wb


# ------------------------------------------------------------------------------------ loss --

def _size_hw(size):
    if isinstance(size, torch.Tensor):
        size = size.tolist()  # the reference does the same host read (PixPro.py:116)
    return int(size[0]), int(size[1])


def add_optical_flow(flow, x_grid, y_grid, size, mask=None):
    """contrast/models/PixPro.py:46-89 -> (out_x, out_y, mask_grid [B,1,G,G] bool or None)."""
    flow = _f32(flow, "flow")
    xg = _f32(x_grid, "x_grid")
    yg = _f32(y_grid, "y_grid")
    B, _, Hin, Win = flow.shape
    H_orig, W_orig = _size_hw(size)
    P = xg[0].numel()
    ox = torch.empty_like(xg)
    oy = torch.empty_like(yg)
    m = _mask_u8(mask, "mask") if mask is not None else None
    mg = torch.empty(xg.shape, device=flow.device, dtype=torch.uint8) if mask is not None else None
    with torch.cuda.device(flow.device):
        _cabi.check(_cabi.lib().pp_add_optical_flow(_ptr(flow), B, Hin, Win, _ptr(xg), _ptr(yg), P, H_orig, W_orig, _ptr(m),
                                                    _div_mode, _ptr(ox), _ptr(oy), _ptr(mg), _stream()),
                    "pp_add_optical_flow")
    if mg is not None:
        mg = mg.view(torch.bool).unsqueeze(1)
    return ox, oy, mg


class _RegressionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, coord_q, coord_k, flow, mask, size, pos_ratio, debug, warped=None):
        q = _f32(q, "q")
        k = _f32(k, "k")
        cq = _f32(coord_q, "coord_q")
        ck = _f32(coord_k, "coord_k")
        B, C, G, G2 = q.shape
        assert G == G2 and k.shape == q.shape and cq.shape == (B, 10) and ck.shape == (B, 10)
        P = G * G
        dev = q.device
        if flow is not None:
            flow = _f32(flow, "flow")
            Hin, Win = flow.shape[-2:]
        else:
            Hin = Win = 0
        m = _mask_u8(mask, "mask") if mask is not None else None
        H_orig, W_orig = size
        L = _cabi.lib()
        ws = torch.empty((L.pp_regression_loss_workspace(B, C, G),), device=dev, dtype=torch.uint8)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        pos_num = torch.empty((B,), device=dev, dtype=torch.float32)
        pos_mean = torch.empty((B,), device=dev, dtype=torch.float32)
        dq = torch.empty_like(q)
        pos_mask = torch.empty((B, P, P), device=dev, dtype=torch.uint8) if debug else None
        centres = torch.empty((4, B, P), device=dev, dtype=torch.float32) if debug else None
        with torch.cuda.device(dev):
            if warped is not None:
                wp = _req(warped, "warped")
                assert flow is None and mask is None and wp.shape == (3, B, P)
                _cabi.check(L.pp_regression_loss_warped(_ptr(q), _ptr(k), B, C, G, _ptr(cq), _ptr(ck), _ptr(wp), H_orig, W_orig,
                                                        float(pos_ratio), _div_mode, _ptr(loss), _ptr(pos_num), _ptr(pos_mean),
                                                        _ptr(dq), _ptr(pos_mask), _ptr(centres), _ptr(ws), _stream()),
                            "pp_regression_loss_warped")
            else:
                _cabi.check(L.pp_regression_loss(_ptr(q), _ptr(k), B, C, G, _ptr(cq), _ptr(ck), _ptr(flow), Hin, Win, _ptr(m),
                                                 H_orig, W_orig, float(pos_ratio), _div_mode, _ptr(loss), _ptr(pos_num),
                                                 _ptr(pos_mean), _ptr(dq), _ptr(pos_mask), _ptr(centres), _ptr(ws), _stream()),
                            "pp_regression_loss")
        ctx.save_for_backward(dq)
        ctx.mark_non_differentiable(pos_num, pos_mean)
        if debug:
            pm = pos_mask.view(torch.bool)
            ctx.mark_non_differentiable(pm, centres)
            return loss, pos_num, pos_mean, pm, centres
        return loss, pos_num, pos_mean

    @staticmethod
    def backward(ctx, g_loss, *unused):
        (dq,) = ctx.saved_tensors
        return (dq * g_loss,) + (None,) * 9


def regression_loss(q, k, coord_q, coord_k, pos_ratio=0.5, flow=None, size=None, mask=None, debug=False, warped=None):
    """contrast/models/PixPro.py:92-247 with the nested-list arguments already unpacked.

    flow may be a dense composite [B,2,H,W] (+ dense mask), or a LazyFlow (+ its LazyMask): the
    sparse correspondence path; or pass `warped` [3,B,P] from sparse_corr() directly (size required).
    Returns (loss, pos_num, pos_mean) — plus (pos_mask [B,P,P] bool, centres [4,B,P]) if debug."""
    if isinstance(flow, LazyFlow):
        if size is None:
            size = tuple(flow.shape[-2:])
        size = _size_hw(size)
        warped = _lazy_warped(flow, mask, coord_q, q.shape[-1], size)
        flow = mask = None
    if warped is not None:
        assert flow is None and mask is None and size is not None, "warped centres replace flow/mask; size is required"
        return _RegressionLoss.apply(q, k.detach(), coord_q, coord_k, None, None, _size_hw(size), pos_ratio, debug, warped)
    if size is None:
        if flow is not None:
            size = tuple(flow.shape[-2:])                          # PixPro.py:121
        else:
            size = (coord_q[0][9].item(), coord_q[0][8].item())    # PixPro.py:123 (same host read)
    size = _size_hw(size)
    return _RegressionLoss.apply(q, k.detach(), coord_q, coord_k, flow, mask, size, pos_ratio, debug)


def _loss_pair_launch(q1, q2, k1, k2, cq1, ck1, cq2, ck2, flow1, flow2, mask1, mask2, size, pos_ratio, warped1, warped2, dqs=None):
    """One pp_regression_loss_pair(_warped) launch; returns (loss [2], pos_num [2,B], pos_mean [2,B], [dq1, dq2])."""
    import ctypes
    qs = [_f32(q1, "q1"), _f32(q2, "q2")]
    ks = [_f32(k1, "k1"), _f32(k2, "k2")]
    cqs = [_f32(cq1, "coord_q1"), _f32(cq2, "coord_q2")]
    cks = [_f32(ck1, "coord_k1"), _f32(ck2, "coord_k2")]
    B, C, G, G2 = qs[0].shape
    assert G == G2 and all(t.shape == qs[0].shape for t in qs + ks) and all(t.shape == (B, 10) for t in cqs + cks)
    dev = qs[0].device
    flows = [None if f is None else _f32(f, "flow") for f in (flow1, flow2)]
    masks = [None if m is None else _mask_u8(m, "mask") for m in (mask1, mask2)]
    Hin = Win = 0
    for f in flows:
        if f is not None:
            Hin, Win = f.shape[-2:]
    H_orig, W_orig = size
    L = _cabi.lib()
    wsz = L.pp_regression_loss_workspace(B, C, G)
    ws = [torch.empty((wsz,), device=dev, dtype=torch.uint8) for _ in range(2)]
    loss = torch.empty((2,), device=dev, dtype=torch.float32)
    pos_num = torch.empty((2, B), device=dev, dtype=torch.float32)
    pos_mean = torch.empty((2, B), device=dev, dtype=torch.float32)
    if dqs is None:
        dqs = [torch.empty_like(qs[0]), torch.empty_like(qs[1])]

    def table(ts):
        return (ctypes.c_void_p * 2)(*[None if t is None else t.data_ptr() for t in ts])

    keep = [table(qs), table(ks), table(cqs), table(cks), table(flows), table(masks), table([loss[0:], loss[1:]]),
            table([pos_num[0], pos_num[1]]), table([pos_mean[0], pos_mean[1]]), table(dqs), table(ws)]
    ptrs = [ctypes.cast(t, ctypes.c_void_p) for t in keep]
    with torch.cuda.device(dev):
        if warped1 is not None:
            wps = [_req(warped1, "warped1"), _req(warped2, "warped2")]
            assert all(f is None for f in flows + masks) and all(t.shape == (3, B, G * G) for t in wps)
            wt = table(wps)
            _cabi.check(L.pp_regression_loss_pair_warped(ptrs[0], ptrs[1], B, C, G, ptrs[2], ptrs[3],
                                                         ctypes.cast(wt, ctypes.c_void_p), H_orig, W_orig, float(pos_ratio),
                                                         _div_mode, ptrs[6], ptrs[7], ptrs[8], ptrs[9], ptrs[10], _stream()),
                        "pp_regression_loss_pair_warped")
        else:
            _cabi.check(L.pp_regression_loss_pair(ptrs[0], ptrs[1], B, C, G, ptrs[2], ptrs[3], ptrs[4], Hin, Win, ptrs[5],
                                                  H_orig, W_orig, float(pos_ratio), _div_mode, ptrs[6], ptrs[7], ptrs[8],
                                                  ptrs[9], ptrs[10], _stream()), "pp_regression_loss_pair")
    return loss, pos_num, pos_mean, dqs


class _RegressionLossPair(torch.autograd.Function):
    """Both directions of the pixel loss (PixPro.py:429-430) in one launch."""

    @staticmethod
    def forward(ctx, q1, q2, k1, k2, cq1, ck1, cq2, ck2, flow1, flow2, mask1, mask2, size, pos_ratio, warped1=None,
                warped2=None):
        loss, pos_num, pos_mean, dqs = _loss_pair_launch(q1, q2, k1, k2, cq1, ck1, cq2, ck2, flow1, flow2, mask1, mask2, size,
                                                         pos_ratio, warped1, warped2)
        ctx.save_for_backward(dqs[0], dqs[1])
        ctx.mark_non_differentiable(pos_num, pos_mean)
        return loss, pos_num, pos_mean

    @staticmethod
    def backward(ctx, g_loss, *unused):
        dq1, dq2 = ctx.saved_tensors
        return (dq1 * g_loss[0], dq2 * g_loss[1]) + (None,) * 14


class _RegressionLossPairJoint(torch.autograd.Function):
    """The same launch for predictions that arrive as ONE tensor q12 = [q1; q2] (both views went through the PPM as one
    batch, PixPro.py:420-430) and a loss that is consumed as the sum of the two directions (`loss = l1 + l2`, :432): the
    gradient leaves as one tensor too.  Against chunk() + _RegressionLossPair + `l[0] + l[1]` this removes eight small
    autograd kernels (select / zeros / copy / add of the 2-vector, two multiplies, the cat of the chunk's backward) from the
    critical path between the loss and the PPM backward."""

    @staticmethod
    def forward(ctx, q12, k1, k2, cq1, ck1, cq2, ck2, flow1, flow2, mask1, mask2, size, pos_ratio, warped1=None, warped2=None):
        q12 = _f32(q12, "q12")
        B = q12.shape[0] // 2
        dq12 = torch.empty_like(q12)
        loss, pos_num, pos_mean, _ = _loss_pair_launch(q12[:B], q12[B:], k1, k2, cq1, ck1, cq2, ck2, flow1, flow2, mask1, mask2,
                                                       size, pos_ratio, warped1, warped2, dqs=[dq12[:B], dq12[B:]])
        ctx.save_for_backward(dq12)
        ctx.mark_non_differentiable(loss, pos_num, pos_mean)
        return loss.sum(), loss, pos_num, pos_mean

    @staticmethod
    def backward(ctx, g_sum, *unused):
        dq12, = ctx.saved_tensors
        return (dq12 * g_sum,) + (None,) * 14


def regression_loss_pair(q1, k1, coord_q1, coord_k1, q2, k2, coord_q2, coord_k2, pos_ratio=0.5, flow1=None, flow2=None,
                         size=None, mask1=None, mask2=None, warped1=None, warped2=None):
    """Two regression_loss calls (the two directions of PixPro.forward) fused into one launch.

    q2=None: q1 is the joint prediction tensor [q1; q2] of shape [2B,C,G,G] (both views went through the PPM as one batch)
    and the result is (loss_1 + loss_2, loss [2], pos_num [2,B], pos_mean [2,B]) with only the sum differentiable — the
    gradient reaches the joint tensor in one piece (_RegressionLossPairJoint).

    flow1/flow2: dense composites (+ dense masks), or the two LazyFlows of one LazyFlowPair (sparse
    correspondence: one pp_sparse_corr launch serves both directions); or pass warped1/warped2 directly.
    Returns (loss [2], pos_num [2,B], pos_mean [2,B]); loss[i] equals regression_loss(q_i, k_i, ...)."""
    if isinstance(flow1, LazyFlow) or isinstance(flow2, LazyFlow):
        if not (isinstance(flow1, LazyFlow) and isinstance(flow2, LazyFlow) and flow1.pair is flow2.pair
                and flow1.direction != flow2.direction):
            raise ValueError("regression_loss_pair: lazy flows must be the two directions of one LazyFlowPair")
        pr = flow1.pair
        for f, m in ((flow1, mask1), (flow2, mask2)):
            if m is not None and not (isinstance(m, LazyMask) and m.pair is pr and m.direction == f.direction):
                raise ValueError("a LazyFlow must come with the LazyMask of the same apply_optical_flow call (or None)")
        if (mask1 is None) != (mask2 is None):
            raise ValueError("regression_loss_pair: lazy masks must be given for both directions or neither")
        if size is None:
            size = tuple(flow1.shape[-2:])
        size = _size_hw(size)
        a1, a2 = (pr.alpha_1, pr.alpha_2) if mask1 is not None else (None, None)
        cf, cb = (coord_q1, coord_q2) if flow1.direction == 0 else (coord_q2, coord_q1)
        wf, wb = sparse_corr(pr.lo_fwd, pr.lo_bwd, cf, cb, q1.shape[-1], size, flow_up=pr.flow_up, alpha_1=a1, alpha_2=a2)
        warped1, warped2 = (wf, wb) if flow1.direction == 0 else (wb, wf)
        flow1 = flow2 = mask1 = mask2 = None
    if warped1 is not None or warped2 is not None:
        assert warped1 is not None and warped2 is not None and size is not None and flow1 is None and flow2 is None
        if q2 is None:
            return _RegressionLossPairJoint.apply(q1, k1.detach(), k2.detach(), coord_q1, coord_k1, coord_q2, coord_k2, None, None,
                                                  None, None, _size_hw(size), pos_ratio, warped1, warped2)
        return _RegressionLossPair.apply(q1, q2, k1.detach(), k2.detach(), coord_q1, coord_k1, coord_q2, coord_k2, None, None,
                                         None, None, _size_hw(size), pos_ratio, warped1, warped2)
    if size is None:
        f = flow1 if flow1 is not None else flow2
        if f is not None:
            size = tuple(f.shape[-2:])
        else:
            size = (coord_q1[0][9].item(), coord_q1[0][8].item())
    size = _size_hw(size)
    if q2 is None:
        return _RegressionLossPairJoint.apply(q1, k1.detach(), k2.detach(), coord_q1, coord_k1, coord_q2, coord_k2, flow1, flow2,
                                              mask1, mask2, size, pos_ratio)
    return _RegressionLossPair.apply(q1, q2, k1.detach(), k2.detach(), coord_q1, coord_k1, coord_q2, coord_k2, flow1, flow2,
                                     mask1, mask2, size, pos_ratio)


# ------------------------------------------------------------------------------------- PPM --

class _PPM(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, feat, val, gamma, clamp_value, final_norm):
        feat = _f32(feat, "feat")
        val = _f32(val, "val")
        assert feat.shape == val.shape and feat.ndim == 4
        B, C, H, W = feat.shape
        P = H * W
        L = _cabi.lib()
        out = torch.empty_like(feat)
        saved = torch.empty((L.pp_ppm_saved_bytes(B, C, P),), device=feat.device, dtype=torch.uint8)
        with torch.cuda.device(feat.device):
            _cabi.check(L.pp_ppm_fwd(_ptr(feat), _ptr(val), B, C, P, float(gamma), float(clamp_value), int(final_norm),
                                     _ptr(out), _ptr(saved), _stream()), "pp_ppm_fwd")
        ctx.save_for_backward(feat, val, out, saved)
        ctx.cfg = (float(gamma), float(clamp_value), int(final_norm))
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        feat, val, out, saved = ctx.saved_tensors
        gamma, cv, final_norm = ctx.cfg
        g = _f32(g, "grad_out")
        B, C, H, W = feat.shape
        P = H * W
        L = _cabi.lib()
        d_feat = torch.empty_like(feat)
        d_val = torch.empty_like(val)
        ws = torch.empty((L.pp_ppm_bwd_workspace(B, C, P),), device=feat.device, dtype=torch.uint8)
        with torch.cuda.device(feat.device):
            _cabi.check(L.pp_ppm_bwd(_ptr(feat), _ptr(val), _ptr(out), _ptr(g), _ptr(saved), B, C, P, gamma, cv, final_norm,
                                     _ptr(d_feat), _ptr(d_val), _ptr(ws), _stream()), "pp_ppm_bwd")
        return d_feat, d_val, None, None, None


def ppm(feat, val, gamma=2.0, clamp_value=0.0, final_norm=True):
    """PixPro.featprop after value_transform (contrast/models/PixPro.py:343-363) fused with the
    caller's F.normalize (:380) when final_norm.  feat, val [B,C,H,W] -> [B,C,H,W]."""
    return _PPM.apply(feat, val, gamma, clamp_value, final_norm)


# -------------------------------------------------------------------------- value transform --

_side_streams = {}


def _side_stream(device, idx=0):
    """Extra streams per device for independent launches inside an op (created lazily)."""
    d = torch.device(device)
    key = (d.index if d.index is not None else torch.cuda.current_device(), idx)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


class _Conv1x1(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, w, bias):
        x = _f32(x, "x")
        w2 = _f32(w, "weight").reshape(w.shape[0], w.shape[1])
        b = _f32(bias, "bias") if bias is not None else None
        B, Cin = x.shape[:2]
        P = x[0, 0].numel()
        Cout = w2.shape[0]
        assert w2.shape[1] == Cin
        y = torch.empty((B, Cout) + tuple(x.shape[2:]), device=x.device, dtype=torch.float32)
        nws = _cabi.lib().pp_conv1x1_fwd_workspace(B, Cin, Cout, P)
        ws = torch.empty((nws,), device=x.device, dtype=torch.uint8) if nws else None
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().pp_conv1x1_fwd(_ptr(x), _ptr(w2), _ptr(b), B, Cin, Cout, P, _ptr(y), _ptr(ws), _stream()),
                        "pp_conv1x1_fwd")
        ctx.save_for_backward(x, w2)
        ctx.has_bias = bias is not None
        ctx.w_shape = tuple(w.shape)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w2 = ctx.saved_tensors
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dx, dw, db = _conv1x1_backward(x, w2, dy, need_x, need_w, need_b)
        return dx, (dw.view(ctx.w_shape) if dw is not None else None), db


def _conv1x1_backward(x, w2, dy, need_x, need_w, need_b):
    """dx, dw, db of the 1x1 convolution (pp_conv1x1_bwd; the three gradients are independent launches)."""
    dy = _f32(dy, "grad_out")
    B, Cin = x.shape[:2]
    P = x[0, 0].numel()
    Cout = w2.shape[0]
    L = _cabi.lib()
    bwd = L.pp_conv1x1_bwd
    dx = torch.empty_like(x) if need_x else None
    dw = torch.empty_like(w2) if need_w else None
    db = torch.empty((Cout,), device=x.device, dtype=torch.float32) if need_b else None
    ws = torch.empty((L.pp_conv1x1_bwd_workspace(B, Cin, Cout, P),), device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        if need_x and (need_w or need_b) and not _serial:
            # dgrad and wgrad (+ bias grad) are independent, latency-bound launches: issue the parameter
            # gradients on a side stream (all buffers were allocated on the current one, which joins below)
            cur = torch.cuda.current_stream(x.device)
            side = _side_stream(x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                _cabi.check(L.pp_conv1x1_bwd(_ptr(x), _ptr(w2), _ptr(dy), B, Cin, Cout, P, None, _ptr(dw), None, _ptr(ws),
                                             _stream()), "pp_conv1x1_bwd")
            if need_b:  # the bias gradient is a third independent launch
                side2 = _side_stream(x.device, 1)
                side2.wait_stream(cur)
                with torch.cuda.stream(side2):
                    _cabi.check(L.pp_conv1x1_bwd(_ptr(x), _ptr(w2), _ptr(dy), B, Cin, Cout, P, None, None, _ptr(db), _ptr(ws),
                                                 _stream()), "pp_conv1x1_bwd")
            _cabi.check(bwd(_ptr(x), _ptr(w2), _ptr(dy), B, Cin, Cout, P, _ptr(dx), None, None, _ptr(ws), _stream()),
                        "pp_conv1x1_bwd")
            cur.wait_stream(side)
            if need_b:
                cur.wait_stream(side2)
        else:
            _cabi.check(bwd(_ptr(x), _ptr(w2), _ptr(dy), B, Cin, Cout, P, _ptr(dx), _ptr(dw), _ptr(db), _ptr(ws), _stream()),
                        "pp_conv1x1_bwd")
    return dx, dw, db


def conv1x1(x, weight, bias=None):
    """1x1 convolution (the PPM value transform, contrast/models/PixPro.py:21-23) on the tcgen05
    tensor cores with 3xTF32 (fp32-accurate).  x [B,Cin,H,W], weight [Cout,Cin,1,1] or [Cout,Cin]."""
    return _Conv1x1.apply(x, weight, bias)


class _FeatProp(torch.autograd.Function):
    """PixPro.featprop with its value transform (contrast/models/PixPro.py:339-363 [+ F.normalize, :380]) as ONE autograd
    node: forward = pp_conv1x1_fwd -> pp_ppm_fwd; backward = pp_ppm_bwd -> pp_conv1x1_bwd, the two gradients of `feat` joined by
    one in-place add.  Same kernels and bits as ppm(feat, conv1x1(feat, w, b)), one node of graph bookkeeping less."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, feat, w, bias, gamma, clamp_value, final_norm):
        feat = _f32(feat, "feat")
        w2 = _f32(w, "weight").reshape(w.shape[0], w.shape[1])
        b = _f32(bias, "bias") if bias is not None else None
        B, C, H, W = feat.shape
        P = H * W
        Cout = w2.shape[0]
        assert w2.shape[1] == C and Cout == C, "the value transform keeps the channel count (PixPro.py:300)"
        L = _cabi.lib()
        val = torch.empty_like(feat)
        out = torch.empty_like(feat)
        nws = L.pp_conv1x1_fwd_workspace(B, C, Cout, P)
        ws = torch.empty((nws,), device=feat.device, dtype=torch.uint8) if nws else None
        saved = torch.empty((L.pp_ppm_saved_bytes(B, C, P),), device=feat.device, dtype=torch.uint8)
        with torch.cuda.device(feat.device):
            _cabi.check(L.pp_conv1x1_fwd(_ptr(feat), _ptr(w2), _ptr(b), B, C, Cout, P, _ptr(val), _ptr(ws), _stream()), "pp_conv1x1_fwd")
            _cabi.check(L.pp_ppm_fwd(_ptr(feat), _ptr(val), B, C, P, float(gamma), float(clamp_value), int(final_norm),
                                     _ptr(out), _ptr(saved), _stream()), "pp_ppm_fwd")
        ctx.save_for_backward(feat, w2, val, out, saved)
        ctx.cfg = (float(gamma), float(clamp_value), int(final_norm))
        ctx.has_bias = bias is not None
        ctx.w_shape = tuple(w.shape)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        feat, w2, val, out, saved = ctx.saved_tensors
        gamma, cv, final_norm = ctx.cfg
        g = _f32(g, "grad_out")
        B, C, H, W = feat.shape
        P = H * W
        L = _cabi.lib()
        d_feat = torch.empty_like(feat)
        d_val = torch.empty_like(val)
        ws = torch.empty((L.pp_ppm_bwd_workspace(B, C, P),), device=feat.device, dtype=torch.uint8)
        with torch.cuda.device(feat.device):
            _cabi.check(L.pp_ppm_bwd(_ptr(feat), _ptr(val), _ptr(out), _ptr(g), _ptr(saved), B, C, P, gamma, cv, final_norm,
                                     _ptr(d_feat), _ptr(d_val), _ptr(ws), _stream()), "pp_ppm_bwd")
        need_w, need_b = ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        # the two gradients of `feat` (similarity branch, value transform) meet in one in-place add.  Adding inside the conv's
        # epilogue instead (read-modify-write) was built and measured: 31 -> 47 us at 7x7 (strided scalar accesses of the
        # thread-staged kernel) and 72 -> 153 us at 28x28 (a dependent global load per epilogue group of the TMA-fed kernel);
        # the separate element-wise pass costs 4 / 20 us.
        dx, dw, db = _conv1x1_backward(feat, w2, d_val, True, need_w, need_b)
        d_feat.add_(dx)
        return d_feat, (dw.view(ctx.w_shape) if dw is not None else None), db, None, None, None


def featprop(feat, weight, bias=None, gamma=2.0, clamp_value=0.0, final_norm=True):
    """ppm(feat, conv1x1(feat, weight, bias), ...) as one autograd node (PixPro.featprop with the single-conv value
    transform, `pixpro_transform_layer=1`); see _FeatProp."""
    return _FeatProp.apply(feat, weight, bias, gamma, clamp_value, final_norm)


# ---------------------------------------------------------------------------- tensor cores --

def tc_gemm_nt(A, B, tma=True, workspace=None):
    """C[b] = A[b] @ B[b].T on the tcgen05 tensor cores with 3xTF32 (fp32-accurate).
    A [batch,M,K], B [batch,N,K] -> [batch,M,N].  tma=True: the TMA-fed warp-specialised kernel (csrc/pp_tc2.cuh; the
    operands are split into hi / lo planes in a workspace first); tma=False: the thread-staged kernel (csrc/pp_tc.cuh)."""
    A = _f32(A, "A")
    B = _f32(B, "B")
    assert A.ndim == 3 and B.ndim == 3 and A.shape[0] == B.shape[0] and A.shape[2] == B.shape[2]
    batch, M, K = A.shape
    N = B.shape[1]
    C = torch.empty((batch, M, N), device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        if tma:
            nbytes = _cabi.lib().pp_tc_gemm_nt_workspace(batch, M, N, K)
            if workspace is None or workspace.numel() < nbytes:
                workspace = torch.empty(nbytes, device=A.device, dtype=torch.uint8)
            _cabi.check(_cabi.lib().pp_tc_gemm_nt_ws(_ptr(A), _ptr(B), _ptr(C), batch, M, N, K, _ptr(workspace), _stream()),
                        "pp_tc_gemm_nt_ws")
        else:
            _cabi.check(_cabi.lib().pp_tc_gemm_nt(_ptr(A), _ptr(B), _ptr(C), batch, M, N, K, _stream()), "pp_tc_gemm_nt")
    return C


def tc_gemm(A, B, a_mn=False, b_mn=False):
    """C[b][m][n] = sum_k A(m,k) B(n,k) on the TMA-fed tcgen05 kernel, operands in place in either layout:
    A [batch,M,K] (or [batch,K,M] with a_mn), B [batch,N,K] (or [batch,K,N] with b_mn) -> [batch,M,N]."""
    A = _f32(A, "A")
    B = _f32(B, "B")
    assert A.ndim == 3 and B.ndim == 3 and A.shape[0] == B.shape[0]
    batch = A.shape[0]
    M, K = (A.shape[2], A.shape[1]) if a_mn else (A.shape[1], A.shape[2])
    N, K2 = (B.shape[2], B.shape[1]) if b_mn else (B.shape[1], B.shape[2])
    assert K == K2
    C = torch.empty((batch, M, N), device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        _cabi.check(_cabi.lib().pp_tc_gemm_ws(_ptr(A), _ptr(B), _ptr(C), batch, M, N, K, int(a_mn), int(b_mn), _stream()), "pp_tc_gemm_ws")
    return C


def tc_gemm_padded(A, B, M, N, K, a_mn=False, b_mn=False, kb=1):
    """pp_tc_gemm_ex: like tc_gemm, but A / B may be padded along their contiguous dimension (the logical extents M, N, K are
    given; the pitch is the tensor's last dimension) and kb > 1 sums the products of kb consecutive batch entries:
    returns [ceil(batch / kb), M, N]."""
    A = _f32(A, "A")
    B = _f32(B, "B")
    assert A.ndim == 3 and B.ndim == 3 and A.shape[0] == B.shape[0] and A.is_contiguous() and B.is_contiguous()
    batch = A.shape[0]
    assert A.shape[1] == (K if a_mn else M) and B.shape[1] == (K if b_mn else N)
    a_pitch, b_pitch = A.shape[2], B.shape[2]
    assert a_pitch >= (M if a_mn else K) and b_pitch >= (N if b_mn else K)
    C = torch.empty(((batch + kb - 1) // kb, M, N), device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        _cabi.check(_cabi.lib().pp_tc_gemm_ex(_ptr(A), _ptr(B), _ptr(C), batch, M, N, K, int(a_mn), int(b_mn), a_pitch, b_pitch,
                                              int(kb), _stream()), "pp_tc_gemm_ex")
    return C


# ---------------------------------------------------------------------- RAFT correlation --

def corr_volume(fmap1, fmap2):
    """CorrBlock.corr (contrast/flow/corr.py:52-60): fmap1, fmap2 [B,D,h,w] -> [B, h*w, h*w] = <f1_i, f2_j> / sqrt(D)."""
    f1, f2 = _f32(fmap1, "fmap1"), _f32(fmap2, "fmap2")
    assert f1.ndim == 4 and f1.shape == f2.shape
    B, D, h, w = f1.shape
    out = torch.empty((B, h * w, h * w), device=f1.device, dtype=torch.float32)
    with torch.cuda.device(f1.device):
        _cabi.check(_cabi.lib().pp_corr_volume(_ptr(f1), _ptr(f2), B, D, h, w, _ptr(out), _stream()), "pp_corr_volume")
    return out


def corr_pool(corr):
    """One pyramid step (corr.py:26-28): avg_pool2d(2, stride=2) over the last two dims of [N,1,h,w] (or [N,h,w])."""
    c = _f32(corr, "corr")
    h, w = c.shape[-2:]
    planes = c.numel() // (h * w)
    out = torch.empty(tuple(c.shape[:-2]) + (h // 2, w // 2), device=c.device, dtype=torch.float32)
    with torch.cuda.device(c.device):
        _cabi.check(_cabi.lib().pp_corr_pool(_ptr(c), planes, h, w, _ptr(out), _stream()), "pp_corr_pool")
    return out


def corr_lookup(pyramid, coords, radius):
    """CorrBlock.__call__ (corr.py:30-50).  pyramid: list of [B*h*w, 1, h>>l, w>>l]; coords [B,2,h,w] -> [B, L*(2r+1)^2, h, w]."""
    import ctypes
    coords = _f32(coords, "coords")
    B, two, h, w = coords.shape
    assert two == 2
    levels = [_f32(p, "pyramid level") for p in pyramid]
    for l, p in enumerate(levels):
        assert p.shape[0] == B * h * w and tuple(p.shape[-2:]) == (h >> l, w >> l), "pyramid level %d has shape %s" % (l, tuple(p.shape))
    K = 2 * radius + 1
    out = torch.empty((B, len(levels) * K * K, h, w), device=coords.device, dtype=torch.float32)
    table = (ctypes.c_void_p * len(levels))(*[p.data_ptr() for p in levels])
    with torch.cuda.device(coords.device):
        _cabi.check(_cabi.lib().pp_corr_lookup(ctypes.cast(table, ctypes.c_void_p), len(levels), _ptr(coords), B, h, w, int(radius),
                                               _div_mode, _ptr(out), _stream()), "pp_corr_lookup")
    return out
