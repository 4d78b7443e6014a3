"""contrast.util — flow chaining / forward-backward consistency utilities of the reference
(contrast/util.py:75-366), backed by the sm_100a kernels.  Signatures and return structures
are the reference's; tensors must live on a CUDA device (the reference hard-codes `.cuda()`
at util.py:196-197 too)."""
import os

import torch

from pixpro_b200 import ops as _ops

from .flow import upflow8


class AverageMeter(object):
    """Computes and stores the average and current value (contrast/util.py:10-29)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


@torch.no_grad()
def normalize_coord(coords):
    """contrast/util.py:334-339"""
    return _ops.normalize_coord(coords)


@torch.no_grad()
def normalize_flow(flow):
    """contrast/util.py:343-348"""
    return _ops.normalize_flow(flow)


@torch.no_grad()
def denormalize_flow(flow_norm):
    """contrast/util.py:352-357"""
    return _ops.denormalize_flow(flow_norm)


@torch.no_grad()
def calc_mask_ratio(mask):
    """contrast/util.py:361-366"""
    if mask is None:
        return None
    if isinstance(mask, _ops.LazyMask):  # sparse mode: logging the ratio is what materialises the dense masks
        mask = mask.dense()
    return _ops.calc_mask_ratio(mask)


@torch.no_grad()
def concat_flow(flows, is_norm=False):
    """contrast/util.py:301-330: chain `num` flow links [num,nb,2,ht,wd] -> [nb,2,ht,wd]."""
    return _ops.concat_flow(flows, is_norm=is_norm)


def all_concat_flow(flow_fwds, flow_bwds, is_norm=False, use_flow_frames=True):
    """contrast/util.py:105-126."""
    if not use_flow_frames:
        return concat_flow(flow_fwds, is_norm), concat_flow(flow_bwds, is_norm)
    num_flow = flow_bwds.shape[0]
    fwd_list, bwd_list = [], []
    for span in range(1, num_flow + 1):  # every contiguous sub-chain, shortest first
        for fwd_s in range(num_flow - span + 1):
            bwd_e = num_flow - fwd_s
            fwd_list.append(concat_flow(flow_fwds[fwd_s:fwd_s + span], is_norm))
            bwd_list.append(concat_flow(flow_bwds[bwd_e - span:bwd_e], is_norm))
    return torch.stack(fwd_list), torch.stack(bwd_list)


@torch.no_grad()
def forward_backward_consistency(flow_fwd, flow_bwd, coords0=None, alpha_1=0.01, alpha_2=0.5, is_norm=False):
    """contrast/util.py:253-297 -> (coords0_norm, coords1_norm, [mask, flow_cycle])."""
    if alpha_1 is None or alpha_2 is None:
        return flow_fwd.clone(), flow_bwd.clone(), [None, None]
    coords1_norm, mask, cycle = _ops.forward_backward_consistency(flow_fwd, flow_bwd, alpha_1, alpha_2, is_norm=is_norm)
    if coords0 is None:
        nb, _, ht, wd = flow_fwd.shape
        ys, xs = torch.meshgrid(torch.arange(ht, device=flow_fwd.device), torch.arange(wd, device=flow_fwd.device),
                                indexing='ij')
        coords0 = torch.stack([xs, ys], dim=0).float().repeat(nb, 1, 1, 1)
        coords0_norm = normalize_coord(coords0)
    else:
        # the reference leaves coords0_norm undefined on this branch (util.py:267-271, a NameError);
        # the only caller never passes coords0
        raise NotImplementedError("forward_backward_consistency: explicit coords0 is not supported by the reference either")
    return coords0_norm, coords1_norm, [mask, cycle]


@torch.no_grad()
def calc_optical_flow(imgs, flow_model, up=False, verbose=False):
    """contrast/util.py:76-103: one estimate per consecutive frame pair, forward in time and backward in time (the backward
    links listed from the last frame to the first) -> two [num_img-1, nb, 2, h, w] tensors.  `up`: the model's up-sampled
    prediction instead of its 1/8-resolution one."""
    assert len(imgs) >= 2
    flow_model.eval()
    pick = 1 if up else 0
    rev = list(imgs)[::-1]
    fwd = torch.stack([flow_model(a, b, upsample=False, test_mode=True)[pick] for a, b in zip(imgs[:-1], imgs[1:])])
    bwd = torch.stack([flow_model(a, b, upsample=False, test_mode=True)[pick] for a, b in zip(rev[:-1], rev[1:])])
    if verbose:
        print("calc_optical_flow: flow_fwds %s flow_bwds %s" % (tuple(fwd.shape), tuple(bwd.shape)))
    return fwd.cuda(), bwd.cuda()


def _estimate_links(orig_imgs, flow_model, args):
    """Links of the whole batch in the loader's layout [B, n, 2, h, w], estimated `flow_bs` samples at a time like
    mem_reduce_calc_optical_flow (util.py:128-171; per-sample results do not depend on the chunking).  Returns
    (fwd, bwd, flow_up): the 1/8-resolution links and flow_up=args.flow_up when the model up-samples bilinearly (RAFT-small:
    the x8 up-sampling is then fused into the chain kernels), the model's own full-resolution prediction and flow_up=False
    when it up-samples by learned convex combination (RAFT-basic)."""
    core = flow_model.module if hasattr(flow_model, 'module') else flow_model
    convex = bool(args.flow_up) and getattr(getattr(core, 'update_block', None), 'mask', None) is not None
    bs = orig_imgs[0].shape[0]
    flow_bs = getattr(args, 'flow_bs', None) or 8
    fs, bs_ = [], []
    for i0 in range(0, bs, flow_bs):
        chunk = [im[i0:i0 + flow_bs].cuda() for im in orig_imgs]
        f, b = calc_optical_flow(chunk, flow_model, up=convex, verbose=bool(getattr(args, 'verbose', False)))
        fs.append(f), bs_.append(b)
    fwd = torch.cat(fs, dim=1).permute(1, 0, 2, 3, 4).contiguous()
    bwd = torch.cat(bs_, dim=1).permute(1, 0, 2, 3, 4).contiguous()
    return fwd, bwd, bool(args.flow_up) and not convex


@torch.no_grad()
def mem_reduce_calc_optical_flow(orig_imgs, flow_model, args):
    """contrast/util.py:128-171 -> chained (flow_fwd, flow_bwd), each [k, nb, 2, H, W] (k = 1, or every sub-chain with
    use_flow_frames)."""
    fwd, bwd, flow_up = _estimate_links(orig_imgs, flow_model, args)
    fwd, bwd = fwd.permute(1, 0, 2, 3, 4), bwd.permute(1, 0, 2, 3, 4)
    if flow_up:
        num, nb, c, h, w = fwd.shape
        fwd = upflow8(fwd.reshape(-1, c, h, w)).reshape(num, nb, c, 8 * h, 8 * w)
        bwd = upflow8(bwd.reshape(-1, c, h, w)).reshape(num, nb, c, 8 * h, 8 * w)
    ff, fb = all_concat_flow(fwd, bwd, is_norm=args.flow_cat_norm, use_flow_frames=args.use_flow_frames and len(orig_imgs) > 2)
    if ff.ndim == 4:
        ff, fb = ff.unsqueeze(0), fb.unsqueeze(0)
    return ff, fb


@torch.no_grad()
def apply_optical_flow(data, flow_model, args):
    """contrast/util.py:175-248.  data follows the loader layout (contrast/data/dataset.py:503):
    data[5] = [target, flow_fwd [B,n,2,h,w], flow_bwd [B,n,2,h,w]], data[6] = [size [B,2], num_img [B,1], ...].
    Returns ([flow_fwd, size, mask_fwd], [flow_bwd, size, mask_bwd])."""
    orig_imgs_tmp = data[6]
    size, num_img = orig_imgs_tmp[0][0], int(orig_imgs_tmp[1][0].item())
    is_mask_flow = args.alpha1 is not None and args.alpha2 is not None
    is_use_flow_frames = args.use_flow_frames and num_img > 2
    if args.use_flow_file:
        _, flow_fwds, flow_bwds = data[5]
        flow_up = args.flow_up
    else:
        # util.py:201-204: links estimated on the fly by the RAFT model (contrast.flow.RAFT: correlation kernels of
        # csrc/pp_corr.cu), then the same stage as for precomputed links
        flow_fwds, flow_bwds, flow_up = _estimate_links(orig_imgs_tmp[2:], flow_model, args)
    debug = bool(getattr(args, 'debug', False))
    sparse = bool(getattr(args, 'flow_sparse', False)) or os.environ.get("PIXPRO_B200_SPARSE", "0") == "1"
    if sparse and not is_use_flow_frames and not debug and not args.flow_cat_norm:
        # sparse correspondence (pp_sparse_corr): nothing is computed here.  The returned LazyFlow / LazyMask
        # objects carry the low-res links; regression_loss evaluates the chain and the FB test only at its
        # G*G grid centres (bit-identical to sampling the dense tensors), and `.dense()` / calc_mask_ratio
        # materialise the dense tensors on demand through the fused path below.
        pair = _ops.LazyFlowPair(flow_fwds.cuda(), flow_bwds.cuda(), flow_up=flow_up,
                                 alpha_1=args.alpha1 if is_mask_flow else None, alpha_2=args.alpha2 if is_mask_flow else None)
        return [pair.flow[0], size, pair.mask[0]], [pair.flow[1], size, pair.mask[1]]
    if not is_use_flow_frames and not debug:
        # fused path: x8 up-sampling, chaining and both FB masks in two launches, nothing else
        # materialised (util.py:185-244 in one pass)
        flow_fwd, flow_bwd, mask_fwd, mask_bwd = _ops.flow_stage(
            flow_fwds.cuda(), flow_bwds.cuda(), flow_up=flow_up,
            alpha_1=args.alpha1 if is_mask_flow else None, alpha_2=args.alpha2 if is_mask_flow else None,
            is_norm=args.flow_cat_norm)
        return [flow_fwd, size, mask_fwd], [flow_bwd, size, mask_bwd]

    # general path (use_flow_frames / debug), composed from the same kernels step by step
    flow_fwds = flow_fwds.cuda().permute(1, 0, 2, 3, 4)
    flow_bwds = flow_bwds.cuda().permute(1, 0, 2, 3, 4)
    if flow_up:
        num, nb, c, h, w = flow_fwds.shape
        flow_fwds = upflow8(flow_fwds.reshape(-1, c, h, w)).reshape(num, nb, c, 8 * h, 8 * w)
        flow_bwds = upflow8(flow_bwds.reshape(-1, c, h, w)).reshape(num, nb, c, 8 * h, 8 * w)
    flow_fwd, flow_bwd = all_concat_flow(flow_fwds, flow_bwds, is_norm=args.flow_cat_norm,
                                         use_flow_frames=is_use_flow_frames)
    if flow_fwd.ndim == 4:
        flow_fwd, flow_bwd = flow_fwd.unsqueeze(0), flow_bwd.unsqueeze(0)
    mask_fwd = mask_bwd = None
    if is_mask_flow:
        mf, mb, cf, cb = [], [], [], []
        for l_fwd, l_bwd in zip(flow_fwd, flow_bwd):
            _, _, (m1, c1) = forward_backward_consistency(l_fwd, l_bwd, alpha_1=args.alpha1, alpha_2=args.alpha2,
                                                          is_norm=args.flow_cat_norm)
            _, _, (m2, c2) = forward_backward_consistency(l_bwd, l_fwd, alpha_1=args.alpha1, alpha_2=args.alpha2,
                                                          is_norm=args.flow_cat_norm)
            mf.append(m1), mb.append(m2), cf.append(c1), cb.append(c2)
        mask_fwd, mask_bwd = torch.stack(mf), torch.stack(mb)
        if debug:
            mask_fwd, mask_bwd = [mask_fwd, torch.stack(cf)], [mask_bwd, torch.stack(cb)]
    if args.flow_cat_norm:
        flow_fwd = torch.stack([denormalize_flow(f) for f in flow_fwd])
        flow_bwd = torch.stack([denormalize_flow(f) for f in flow_bwd])
    if not is_use_flow_frames:
        flow_fwd, flow_bwd = flow_fwd[-1], flow_bwd[-1]
        if mask_fwd is None or mask_bwd is None:
            if debug:
                mask_fwd, mask_bwd = [None, None], [None, None]
        elif isinstance(mask_fwd, list):
            mask_fwd, mask_bwd = [m[-1] for m in mask_fwd], [m[-1] for m in mask_bwd]
        else:
            mask_fwd, mask_bwd = mask_fwd[-1], mask_bwd[-1]
    return [flow_fwd, size, mask_fwd], [flow_bwd, size, mask_bwd]


# ------------------------------------------------------------------------------------------------
# Names of the reference's contrast/util.py that are NOT on the pixel path (AverageMeter,
# MyHelpFormatter, dist_collect, reduce_tensor, ...) are served from the reference's own file when the
# reference tree is on sys.path behind this package (see contrast/__init__.py), so that the reference's
# unmirrored modules (`from contrast.util import MyHelpFormatter` in contrast/option.py) keep working.
_reference_util = None


def __getattr__(name):
    global _reference_util
    if name.startswith("__"):
        raise AttributeError(name)
    if _reference_util is None:
        import importlib.util
        import os
        import contrast as _pkg
        here = os.path.dirname(os.path.abspath(__file__))
        for d in list(_pkg.__path__):
            cand = os.path.join(d, "util.py")
            if os.path.abspath(d) != here and os.path.isfile(cand):
                spec = importlib.util.spec_from_file_location("contrast._reference_util", cand)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _reference_util = mod
                break
        else:
            _reference_util = False
    if _reference_util and hasattr(_reference_util, name):
        return getattr(_reference_util, name)
    raise AttributeError(f"module 'contrast.util' has no attribute {name!r} "
                         "(not on the pixel path; put the reference tree on sys.path after this package to get its own)")
