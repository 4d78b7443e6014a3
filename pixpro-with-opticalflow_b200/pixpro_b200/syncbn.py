"""Synchronised batch normalisation in three launches and ONE collective per direction (csrc/pp_syncbn.cu).

The reference converts every BatchNorm of encoder / projector to `torch.nn.SyncBatchNorm`
(contrast/models/PixPro.py:289-292, 315-317) and trains under DDP (main_pretrain.py:78).  torch's implementation
issues ~10 small launches and an all_gather per layer and direction from a Python autograd.Function; with 318 layer
calls per step the multi-GPU step is bound by the host's issue rate (profiles/r01_s3_pretrain_ddp.txt).
`FastSyncBatchNorm` is a subclass (same parameters, buffers, `state_dict` keys, eval behaviour and
`convert_sync_batchnorm` compatibility) whose training forward / backward are:

    forward : pp_bn_stats -> all_gather of one [2C+1] vector (local mean, M2, count) -> pp_bn_apply
    backward: pp_bn_bwd_stats -> all_reduce(SUM) of one [2C] vector -> pp_bn_bwd_apply

fp32 or bf16 activations, NCHW-contiguous or channels_last, fp32 statistics.  Everything else (eval mode, fp16 or
non-CUDA inputs, `track_running_stats=False` in eval) goes through the parent class.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _cabi

_LAYOUT_NCHW, _LAYOUT_NHWC = 0, 1
_DTYPES = {torch.float32: 0, torch.bfloat16: 1}


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _layout_of(x):
    """(tensor in a supported memory layout, layout id, N, C, HW)"""
    N, C = x.shape[0], x.shape[1]
    HW = 1
    for d in x.shape[2:]:
        HW *= d
    if x.dim() == 4 and HW > 1 and C > 1 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        return x, _LAYOUT_NHWC, N, C, HW
    return x.contiguous(), _LAYOUT_NCHW, N, C, HW


# --- the four kernel calls (module-level so that host-logic tests can substitute torch restatements on CPU) ---------------

def bn_stats(x, layout, N, C, HW, ws):
    stats = torch.empty(2 * C + 1, device=x.device, dtype=torch.float32)
    _cabi.check(_cabi.lib().pp_bn_stats(_ptr(x), N, C, HW, layout, _DTYPES[x.dtype], _ptr(ws), _ptr(stats), _stream()), "pp_bn_stats")
    return stats


def bn_apply(x, layout, N, C, HW, stats, nranks, weight, bias, running_mean, running_var, eps, momentum):
    """stats: [nranks, 2C+1] gathered rows.  Returns y and save = [mean (C) | invstd (C) | total count (1)]."""
    y = torch.empty_like(x)
    save = torch.empty(2 * C + 1, device=x.device, dtype=torch.float32)
    _cabi.check(_cabi.lib().pp_bn_apply(_ptr(x), _ptr(y), N, C, HW, layout, _DTYPES[x.dtype], _ptr(stats), nranks, _ptr(weight), _ptr(bias),
                                        _ptr(running_mean), _ptr(running_var), float(eps), float(momentum), _ptr(save), _ptr(save[C:]),
                                        _stream()), "pp_bn_apply")
    return y, save


def bn_bwd_stats(dy, x, layout, N, C, HW, save, ws, need_w, need_b):
    sums = torch.empty(2 * C, device=x.device, dtype=torch.float32)
    gw = torch.empty(C, device=x.device, dtype=torch.float32) if need_w else None
    gb = torch.empty(C, device=x.device, dtype=torch.float32) if need_b else None
    _cabi.check(_cabi.lib().pp_bn_bwd_stats(_ptr(dy), _ptr(x), N, C, HW, layout, _DTYPES[x.dtype], _ptr(save), _ptr(save[C:]), _ptr(ws),
                                            _ptr(sums), _ptr(gw), _ptr(gb), _stream()), "pp_bn_bwd_stats")
    return sums, gw, gb


def bn_bwd_apply(dy, x, layout, N, C, HW, save, weight, sums):
    dx = torch.empty_like(x)
    _cabi.check(_cabi.lib().pp_bn_bwd_apply(_ptr(dy), _ptr(x), _ptr(dx), N, C, HW, layout, _DTYPES[x.dtype], _ptr(save), _ptr(save[C:]),
                                            _ptr(weight), _ptr(sums), _ptr(save[2 * C:]), _stream()), "pp_bn_bwd_apply")
    return dx


def _dev_ctx(x):
    import contextlib
    return torch.cuda.device(x.device) if x.is_cuda else contextlib.nullcontext()


def _world(group):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _all_reduce(t, group):
    if _world(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def _all_gather(t, group):
    """[S] -> ([world, S], world)"""
    world = _world(group)
    if world == 1:
        return t, 1
    out = torch.empty(world * t.numel(), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t, group=group)
    return out, world


class _SyncBNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group, ws):
        x, layout, N, C, HW = _layout_of(x)
        with _dev_ctx(x):
            stats, nranks = _all_gather(bn_stats(x, layout, N, C, HW, ws), group)
            y, save = bn_apply(x, layout, N, C, HW, stats, nranks, weight, bias, running_mean, running_var, eps, momentum)
        ctx.save_for_backward(x, weight, save)
        ctx.cfg = (layout, N, C, HW, group, ws)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, save = ctx.saved_tensors
        layout, N, C, HW, group, ws = ctx.cfg
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dy = dy.contiguous(memory_format=torch.channels_last) if layout == _LAYOUT_NHWC else dy.contiguous()
        need_w = weight is not None and ctx.needs_input_grad[1]
        need_b = ctx.needs_input_grad[2]
        with _dev_ctx(x):
            sums, gw, gb = bn_bwd_stats(dy, x, layout, N, C, HW, save, ws, need_w, need_b)
            dx = None
            if ctx.needs_input_grad[0]:
                _all_reduce(sums, group)
                dx = bn_bwd_apply(dy, x, layout, N, C, HW, save, weight, sums)
        if gw is not None and weight is not None:
            gw = gw.to(weight.dtype)
        return dx, gw, gb, None, None, None, None, None, None


class FastSyncBatchNorm(nn.SyncBatchNorm):
    """Drop-in for torch.nn.SyncBatchNorm (see the module docstring)."""

    def _workspace(self, device):
        ws = getattr(self, "_pp_ws", None)
        if ws is None or ws.device != device:
            ws = torch.zeros(_cabi.lib().pp_bn_workspace(self.num_features), device=device, dtype=torch.uint8)
            self._pp_ws = ws  # plain attribute: not a buffer, not in the state_dict
        return ws

    def forward(self, input):
        fast = (self.training and input.is_cuda and input.dtype in _DTYPES and input.dim() >= 2 and
                (self.weight is None or self.weight.dtype == torch.float32))
        if not fast:
            return super().forward(input)
        # torch.nn.modules.batchnorm: bookkeeping of the exponential average factor
        factor = 0.0 if self.momentum is None else self.momentum
        running_mean = running_var = None
        if self.track_running_stats:
            running_mean, running_var = self.running_mean, self.running_var
            if self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
                if self.momentum is None:
                    factor = 1.0 / float(self.num_batches_tracked)
        return _SyncBNFunction.apply(input, self.weight, self.bias, running_mean, running_var, self.eps, factor, self.process_group,
                                     self._workspace(input.device))


def convert_fast_sync_batchnorm(module, process_group=None):
    """Replaces every BatchNorm / SyncBatchNorm below `module` (in place for containers, like
    nn.SyncBatchNorm.convert_sync_batchnorm) by a FastSyncBatchNorm sharing its parameters and buffers."""
    out = module
    if isinstance(module, nn.modules.batchnorm._BatchNorm) and not isinstance(module, FastSyncBatchNorm):
        out = FastSyncBatchNorm(module.num_features, module.eps, module.momentum, module.affine, module.track_running_stats,
                                process_group if process_group is not None else getattr(module, "process_group", None))
        if module.affine:
            with torch.no_grad():
                out.weight = module.weight
                out.bias = module.bias
        out.running_mean = module.running_mean
        out.running_var = module.running_var
        out.num_batches_tracked = module.num_batches_tracked
        out.training = module.training
        if hasattr(module, "qconfig"):
            out.qconfig = module.qconfig
    for name, child in module.named_children():
        out.add_module(name, convert_fast_sync_batchnorm(child, process_group))
    return out
