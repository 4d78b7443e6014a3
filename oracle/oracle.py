"""numpy front-end of the CPU oracle (oracle/pixpro_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of pixpro_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.

Parity status: pinned against the reference's PyTorch implementation run on CPU
(oracle/pin_against_reference.py -> tests/golden/).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpixpro_oracle.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile the C restatement with gcc (-ffp-contract=off is mandatory)."""
    src = os.path.join(_HERE, "pixpro_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libpixpro_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_fb_alpha2_eff.restype = ctypes.c_float
        _lib.orc_fb_alpha2_eff.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_int]
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_u8p)


def _out_f(shape):
    a = np.empty(shape, np.float32)
    return a, a.ctypes.data_as(_f32p)


def _out_u8(shape):
    a = np.empty(shape, np.uint8)
    return a, a.ctypes.data_as(_u8p)


_L = ctypes.c_long
_I = ctypes.c_int
_D = ctypes.c_double


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(_I(int(n)))


def upflow8(flow):
    """contrast/flow/utils/utils.py:87-89.  flow [N,2,h,w] -> [N,2,8h,8w]."""
    flow, fp = _f(flow)
    N, C, h, w = flow.shape
    out, op = _out_f((N, C, 8 * h, 8 * w))
    lib().orc_upflow8(fp, _L(N * C), _I(h), _I(w), op)
    return out


def _normalize(x, kind, div_mode):
    x, xp = _f(x)
    B, C, H, W = x.shape
    assert C == 2
    out, op = _out_f(x.shape)
    lib().orc_normalize(xp, _L(B), _I(H), _I(W), _I(kind), _I(div_mode), op)
    return out


def normalize_coord(c, div_mode=0):
    """contrast/util.py:334-339"""
    return _normalize(c, 0, div_mode)


def normalize_flow(f, div_mode=0):
    """contrast/util.py:343-348"""
    return _normalize(f, 1, div_mode)


def denormalize_flow(f, div_mode=0):
    """contrast/util.py:352-357"""
    return _normalize(f, 2, div_mode)


def grid_sample_bilinear(inp, grid):
    """F.grid_sample(inp, grid, mode='bilinear', padding_mode='zeros', align_corners=True)"""
    inp, ip = _f(inp)
    grid, gp = _f(grid)
    N, C, H, W = inp.shape
    _, Ho, Wo, _ = grid.shape
    out, op = _out_f((N, C, Ho, Wo))
    lib().orc_grid_sample_bilinear(ip, _L(N), _I(C), _I(H), _I(W), gp, _I(Ho), _I(Wo), op)
    return out


def concat_flow(flows, is_norm=False, div_mode=0):
    """contrast/util.py:301-330.  flows [n,B,2,H,W] -> [B,2,H,W]."""
    flows, fp = _f(flows)
    n, B, C, H, W = flows.shape
    assert C == 2
    out, op = _out_f((B, 2, H, W))
    HW = H * W
    lib().orc_concat_flow(fp, _I(n), _L(B), _I(H), _I(W), _L(B * 2 * HW), _L(2 * HW),
                          _I(int(is_norm)), _I(div_mode), op)
    return out


def forward_backward_consistency(fwd, bwd, alpha_1=0.01, alpha_2=0.5, is_norm=False, div_mode=0):
    """contrast/util.py:253-297 -> (coords1_norm, mask bool [B,H,W], cycle [B,2,H,W])."""
    fwd, fp = _f(fwd)
    bwd, bp = _f(bwd)
    B, _, H, W = fwd.shape
    mask, mp = _out_u8((B, H, W))
    cyc, cp = _out_f((B, 2, H, W))
    c1, c1p = _out_f((B, 2, H, W))
    lib().orc_fb_consistency(fp, bp, _L(B), _I(H), _I(W), _D(alpha_1), _D(alpha_2),
                             _I(int(is_norm)), _I(div_mode), mp, cp, c1p)
    return c1, mask.astype(bool), cyc


def fb_alpha2_eff(alpha_2, H, W):
    return float(lib().orc_fb_alpha2_eff(_D(alpha_2), _I(H), _I(W)))


def calc_mask_ratio(mask):
    """contrast/util.py:361-366"""
    mask, mp = _u8(mask)
    B, H, W = mask.shape
    out, op = _out_f((B,))
    lib().orc_calc_mask_ratio(mp, _L(B), _I(H), _I(W), op)
    return out


def add_optical_flow(flow, x_grid, y_grid, size, mask=None, div_mode=0):
    """contrast/models/PixPro.py:46-89 -> (out_x, out_y, mask_grid or None); grids [B,G,G]."""
    flow, fp = _f(flow)
    x_grid, xp = _f(x_grid)
    y_grid, yp = _f(y_grid)
    B, _, Hin, Win = flow.shape
    P = int(np.prod(x_grid.shape[1:]))
    ox, oxp = _out_f(x_grid.shape)
    oy, oyp = _out_f(x_grid.shape)
    if mask is not None:
        mask, mp = _u8(mask)
        mg, mgp = _out_u8(x_grid.shape)
    else:
        mp, mg, mgp = None, None, None
    lib().orc_add_optical_flow(fp, _L(B), _I(Hin), _I(Win), xp, yp, _I(P), _I(int(size[0])), _I(int(size[1])),
                               mp, _I(div_mode), oxp, oyp, mgp)
    return ox, oy, (mg.astype(bool) if mg is not None else None)


def regression_loss(q, k, coord_q, coord_k, pos_ratio=0.5, flow=None, size=None, mask=None,
                    div_mode=0, want_grad=True):
    """contrast/models/PixPro.py:92-247.

    q,k [B,C,G,G]; coord_* [B,10]; flow [B,2,Hin,Win] (the q-side flow) or None; size (H_orig,W_orig)
    (defaults: flow's shape, or coord[0][9], coord[0][8] on the no-flow path, :117-123);
    mask bool [B,Hin,Win] or None.
    Returns dict(loss, pos_num, pos_mean, pos_mask[B,P,P] bool, cqx,cqy,ckx,cky [B,P], dq [B,C,G,G]).
    """
    q, qp = _f(q)
    k, kp = _f(k)
    coord_q, cqp = _f(coord_q)
    coord_k, ckp = _f(coord_k)
    B, C, G, G2 = q.shape
    assert G == G2
    P = G * G
    if flow is not None:
        flow, fp = _f(flow)
        Hin, Win = flow.shape[-2:]
        if size is None:
            size = (Hin, Win)
    else:
        fp, Hin, Win = None, 0, 0
        if size is None:
            size = (int(coord_q[0][9]), int(coord_q[0][8]))
    if mask is not None:
        mask, mp = _u8(mask)
    else:
        mp = None
    loss, lp = _out_f((1,))
    pn, pnp = _out_f((B,))
    pm, pmp = _out_f((B,))
    pmask, pmaskp = _out_u8((B, P, P))
    cqx, cqxp = _out_f((B, P))
    cqy, cqyp = _out_f((B, P))
    ckx, ckxp = _out_f((B, P))
    cky, ckyp = _out_f((B, P))
    if want_grad:
        dq, dqp = _out_f(q.shape)
    else:
        dq, dqp = None, None
    lib().orc_regression_loss(qp, kp, _L(B), _I(C), _I(G), cqp, ckp, fp, _I(Hin), _I(Win), mp,
                              _I(int(size[0])), _I(int(size[1])), _D(pos_ratio), _I(div_mode),
                              lp, pnp, pmp, pmaskp, cqxp, cqyp, ckxp, ckyp, dqp)
    return dict(loss=float(loss[0]), pos_num=pn, pos_mean=pm, pos_mask=pmask.astype(bool),
                cqx=cqx, cqy=cqy, ckx=ckx, cky=cky, dq=dq)


def near_threshold_pairs(o, coord_q, coord_k, G, size, pos_ratio, tol=1e-5):
    """Number of (i, j) pairs per sample whose normalised centre distance lies within `tol` of pos_ratio
    (SURVEY.md §8d asks for this count beside every positive-mask comparison: such pairs are the ones a 1-ulp
    difference in a coordinate could flip).  o: the dict regression_loss() returned; distances are recomputed in
    fp32 with the reference's formula (PixPro.py:140-157,217-218)."""
    cq = np.asarray(coord_q, np.float32)
    ck = np.asarray(coord_k, np.float32)
    H, W = np.float32(size[0] - 1), np.float32(size[1] - 1)
    g = np.float32(G)
    qd = np.sqrt((((cq[:, 2] - cq[:, 0]) / g) * W) ** 2 + (((cq[:, 3] - cq[:, 1]) / g) * H) ** 2)
    kd = np.sqrt((((ck[:, 2] - ck[:, 0]) / g) * W) ** 2 + (((ck[:, 3] - ck[:, 1]) / g) * H) ** 2)
    md = np.maximum(qd, kd).astype(np.float32)[:, None, None]
    dx = o["cqx"][:, :, None] - o["ckx"][:, None, :]
    dy = o["cqy"][:, :, None] - o["cky"][:, None, :]
    dist = np.sqrt(dx * dx + dy * dy).astype(np.float32) / md
    return (np.abs(dist - np.float32(pos_ratio)) < tol).reshape(dist.shape[0], -1).sum(1)


def featprop(feat, val, gamma=2.0, clamp_value=0.0, final_norm=True):
    """contrast/models/PixPro.py:339-363 (+ F.normalize at :380 when final_norm).

    feat [B,C,G,G]; val = value_transform(feat) [B,C,G,G]."""
    feat, fp = _f(feat)
    val, vp = _f(val)
    B, C = feat.shape[:2]
    P = int(np.prod(feat.shape[2:]))
    out, op = _out_f(feat.shape)
    lib().orc_featprop(fp, vp, _L(B), _I(C), _I(P), _D(gamma), _D(clamp_value), _I(int(final_norm)), op)
    return out


def featprop_bwd(feat, val, g, gamma=2.0, clamp_value=0.0, final_norm=True):
    """Backward of featprop: returns (d_feat through the similarity, d_val)."""
    feat, fp = _f(feat)
    val, vp = _f(val)
    g, gp = _f(g)
    B, C = feat.shape[:2]
    P = int(np.prod(feat.shape[2:]))
    df, dfp = _out_f(feat.shape)
    dv, dvp = _out_f(feat.shape)
    lib().orc_featprop_bwd(fp, vp, gp, _L(B), _I(C), _I(P), _D(gamma), _D(clamp_value),
                           _I(int(final_norm)), dfp, dvp)
    return df, dv


def flow_stage(lo_fwd, lo_bwd, flow_up=True, alpha_1=0.01, alpha_2=0.5, is_norm=False, div_mode=0):
    """Flow stage of contrast/util.py:175-248 (use_flow_file, not use_flow_frames).

    lo_fwd, lo_bwd: loader layout [B,n,2,h,w].  Returns (flow_fwd, flow_bwd [B,2,H,W],
    mask_fwd, mask_bwd bool [B,H,W] or None)."""
    lo_fwd, fp = _f(lo_fwd)
    lo_bwd, bp = _f(lo_bwd)
    B, n, _, h, w = lo_fwd.shape
    H, W = (8 * h, 8 * w) if flow_up else (h, w)
    ff, ffp = _out_f((B, 2, H, W))
    fb, fbp = _out_f((B, 2, H, W))
    use_mask = alpha_1 is not None and alpha_2 is not None
    if use_mask:
        mf, mfp = _out_u8((B, H, W))
        mb, mbp = _out_u8((B, H, W))
    else:
        mf = mb = mfp = mbp = None
    lib().orc_flow_stage(fp, bp, _L(B), _I(n), _I(h), _I(w), _I(int(flow_up)), _I(int(use_mask)),
                         _D(alpha_1 or 0.0), _D(alpha_2 or 0.0), _I(int(is_norm)), _I(div_mode),
                         ffp, fbp, mfp, mbp)
    return ff, fb, (mf.astype(bool) if use_mask else None), (mb.astype(bool) if use_mask else None)


# ------------------------------------------------------------------ optimizer side


def sparse_corr(lo_fwd, lo_bwd, coord_fwd, coord_bwd, grid, size, flow_up=True, alpha_1=0.01, alpha_2=0.5, div_mode=0):
    """Flow stage + add_optical_flow evaluated only at the grid centres (orc_sparse_corr): the restatement of the
    sparse correspondence mode.  Returns (warped_fwd, warped_bwd), each [3,B,P] = (x, y, mask bit) or None."""
    lo_fwd, fp = _f(lo_fwd)
    lo_bwd, bp = _f(lo_bwd)
    B, n, _, h, w = lo_fwd.shape
    use_mask = alpha_1 is not None and alpha_2 is not None
    P = grid * grid
    cf = cb = cfp = cbp = wf = wb = wfp = wbp = None
    if coord_fwd is not None:
        cf, cfp = _f(coord_fwd)
        wf, wfp = _out_f((3, B, P))
    if coord_bwd is not None:
        cb, cbp = _f(coord_bwd)
        wb, wbp = _out_f((3, B, P))
    lib().orc_sparse_corr(fp, bp, _L(B), _I(n), _I(h), _I(w), _I(int(flow_up)), _I(int(use_mask)), _D(alpha_1 or 0.0),
                          _D(alpha_2 or 0.0), cfp, cbp, _I(grid), _I(int(size[0])), _I(int(size[1])), _I(div_mode), wfp, wbp)
    return wf, wb

def ema_update(k, q, m, one_minus_m=None):
    """contrast/models/PixPro.py:330-331, in place on a copy of k; returns the new k."""
    k, kp = _f(np.array(k, dtype=np.float32, copy=True))
    q, qp = _f(q)
    L = lib()
    L.orc_ema_update.argtypes = [_f32p, _f32p, _L, _D, _D]
    L.orc_ema_update(kp, qp, k.size, float(m), float(1.0 - m if one_minus_m is None else one_minus_m))
    return k


def lars_sgd_step(p, g, buf, wd, lr, mom, damp=0.0, lars=True, first=False, trust=0.001, eps=1e-8):
    """One tensor of LARS.step() around SGD (contrast/lars.py:109-152).  Returns (p, buf, rate)."""
    p, pp_ = _f(np.array(p, dtype=np.float32, copy=True))
    g, gp = _f(g)
    buf, bp = _f(np.array(buf if buf is not None else np.zeros_like(p), dtype=np.float32, copy=True))
    L = lib()
    L.orc_lars_sgd_step.restype = ctypes.c_float
    L.orc_lars_sgd_step.argtypes = [_f32p, _f32p, _f32p, _L, _D, _D, _D, _D, _I, _I, _D, _D]
    rate = L.orc_lars_sgd_step(pp_, gp, bp, p.size, float(wd), float(lr), float(mom), float(damp), int(lars), int(first),
                               float(trust), float(eps))
    return p, buf, float(rate)


# ---------------------------------------------------------------- RAFT correlation

def corr_volume(fmap1, fmap2):
    """contrast/flow/corr.py:52-60.  fmap1, fmap2 [B,D,h,w] -> [B, h*w, h*w]."""
    f1, p1 = _f(fmap1)
    f2, p2 = _f(fmap2)
    B, D, h, w = f1.shape
    out, op = _out_f((B, h * w, h * w))
    lib().orc_corr_volume(p1, p2, _L(B), _I(D), _I(h), _I(w), op)
    return out


def corr_pool(corr):
    """contrast/flow/corr.py:26-28: avg_pool2d(2, stride 2) over the last two dims."""
    c, cp = _f(corr)
    h, w = c.shape[-2:]
    planes = c.size // (h * w)
    out, op = _out_f(tuple(c.shape[:-2]) + (h // 2, w // 2))
    lib().orc_corr_pool(cp, _L(planes), _I(h), _I(w), op)
    return out


def corr_lookup(pyramid, coords, radius, div_mode=0):
    """contrast/flow/corr.py:30-50.  pyramid: list of [B*h*w, 1, h>>l, w>>l]; coords [B,2,h,w] -> [B, L*(2r+1)^2, h, w]."""
    import ctypes
    coords, cp = _f(coords)
    B, _, h, w = coords.shape
    keep = [_f(p) for p in pyramid]
    table = (ctypes.c_void_p * len(keep))(*[ctypes.cast(pp_, ctypes.c_void_p).value for _, pp_ in keep])
    K = 2 * radius + 1
    out, op = _out_f((B, len(keep) * K * K, h, w))
    lib().orc_corr_lookup(table, _I(len(keep)), cp, _L(B), _I(h), _I(w), _I(radius), _I(div_mode), op)
    return out
