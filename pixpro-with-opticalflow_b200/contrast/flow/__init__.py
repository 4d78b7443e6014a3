"""contrast.flow — only `upflow8` is on the hot path (contrast/flow/__init__.py:2 re-exports
it from flow/utils/utils.py:87-89).  The RAFT estimator itself is out of scope: flows are
precomputed (`--use_flow_file`), SURVEY.md §2.1 row 6."""
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


__all__ = ['upflow8']
