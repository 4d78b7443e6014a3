"""contrast.flow — the reference package's public names (contrast/flow/__init__.py:1-3): `RAFT` (raft.py), `upflow8`
(flow/utils/utils.py:87-89) and `InputPadder` (:7-24).  The estimator's convolutions are plain PyTorch / cuDNN; its
correlation volume, pyramid, lookup and the x8 up-sampling run on this package's kernels (corr.py, raft.py)."""
import torch.nn.functional as F

from pixpro_b200 import ops as _ops


def upflow8(flow, mode='bilinear'):
    """8 * F.interpolate(flow, (8h, 8w), mode, align_corners=True) — flow/utils/utils.py:87-89."""
    if mode != 'bilinear':
        # only the bilinear mode is used on the hot path (util.py:187-188); other modes are
        # a plain library call, exactly as in the reference
        new_size = (8 * flow.shape[2], 8 * flow.shape[3])
        return 8 * F.interpolate(flow, size=new_size, mode=mode, align_corners=True)
    return _ops.upflow8(flow)


class InputPadder:
    """flow/utils/utils.py:7-24: replicate-pads images so that both sides are multiples of 8.  The horizontal padding is
    split left / right; the vertical one is split top / bottom in 'sintel' mode and goes to the bottom otherwise."""

    def __init__(self, dims, mode='sintel'):
        self.ht, self.wd = dims[-2:]
        extra_h, extra_w = (-self.ht) % 8, (-self.wd) % 8
        top = extra_h // 2 if mode == 'sintel' else 0
        self._pad = [extra_w // 2, extra_w - extra_w // 2, top, extra_h - top]  # F.pad order: left, right, top, bottom

    def pad(self, *inputs):
        return [F.pad(x, self._pad, mode='replicate') for x in inputs]

    def unpad(self, x):
        left, right, top, bottom = self._pad
        return x[..., top:x.shape[-2] - bottom, left:x.shape[-1] - right]


from .raft import RAFT  # noqa: E402  (after upflow8: raft.py imports nothing from here, kept last like the reference's order)

__all__ = ['RAFT', 'InputPadder', 'upflow8']
