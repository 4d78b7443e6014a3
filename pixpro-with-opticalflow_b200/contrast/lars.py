"""`contrast.lars` of the reference (contrast/lars.py) on the fused multi-tensor kernels.

Same public surface: `add_weight_decay(model, weight_decay, skip_list)` and the optimizer wrapper
`LARS(optimizer, eps=1e-8, trust_coef=0.001)` with `step / zero_grad / state_dict / load_state_dict /
param_groups / state / add_param_group`.  `step()` runs the whole update — weight decay folded into the
gradient, per-parameter norms, adaptive rate, SGD momentum, parameter update — as three kernel launches
over all parameters and without the reference's per-parameter host synchronisations
(`if param_norm > 0 and grad_norm > 0`, lars.py:132).  The wrapped optimizer must be `torch.optim.SGD`
(what main_pretrain.py builds); its state (`momentum_buffer`) stays where torch keeps it, so
checkpoints are interchangeable.  One deliberate difference: the reference rewrites `p.grad` on the way
(weight decay added, then scaled); here gradients are left untouched.
"""
import torch
from torch.optim.optimizer import Optimizer

from pixpro_b200 import _cabi
from pixpro_b200.optim import LarsSgdStep

__all__ = ['LARS']


def add_weight_decay(model, weight_decay=1e-5, skip_list=()):
    """Two parameter groups (lars.py:7-31): 1-D parameters (biases, norm scales) and `skip_list` names get no
    weight decay and are ignored by LARS; everything else decays and is LARS-scaled."""
    groups = {True: [], False: []}
    for name, param in model.named_parameters():
        if param.requires_grad:
            groups[param.ndim == 1 or name in skip_list].append(param)
    return [{'params': groups[True], 'weight_decay': 0, 'ignore': True},
            {'params': groups[False], 'weight_decay': weight_decay, 'ignore': False}]


class LARS(Optimizer):
    def __init__(self, optimizer, eps=1e-8, trust_coef=0.001):
        if eps < 0.0:
            raise ValueError('invalid epsilon value: , %f' % eps)
        if trust_coef < 0.0:
            raise ValueError("invalid trust coefficient: %f" % trust_coef)
        if not isinstance(optimizer, torch.optim.SGD):
            raise NotImplementedError("LARS (B200): the wrapped optimizer must be torch.optim.SGD")
        self.optim = optimizer
        self.eps = eps
        self.trust_coef = trust_coef
        self._fused = LarsSgdStep()

    def __getstate__(self):
        return (self.optim, {'eps': self.eps, 'trust_coef': self.trust_coef})

    def __setstate__(self, state):
        self.optim, d = state
        self.eps, self.trust_coef = d['eps'], d['trust_coef']
        self._fused = LarsSgdStep()

    def __repr__(self):
        return '%s(%r)' % (self.__class__.__name__, self.optim)

    @property
    def param_groups(self):
        return self.optim.param_groups

    @property
    def state(self):
        return self.optim.state

    def state_dict(self):
        return self.optim.state_dict()

    def load_state_dict(self, state_dict):
        self.optim.load_state_dict(state_dict)

    def zero_grad(self, *args, **kwargs):
        self.optim.zero_grad(*args, **kwargs)

    def add_param_group(self, param_group):
        self.optim.add_param_group(param_group)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries = []
        for group in self.optim.param_groups:
            if group.get('nesterov') or group.get('maximize'):
                raise NotImplementedError("LARS (B200): nesterov / maximize are not supported")
            ignore = group.get('ignore', None)
            lars = ignore is not None and not ignore  # lars.py:123
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise _cabi.PixProB200Error("LARS (B200): sparse gradients are not supported")
                st = self.optim.state[p]
                buf, first = st.get('momentum_buffer'), False
                if group['momentum'] != 0 and buf is None:
                    buf = st['momentum_buffer'] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    first = True
                entries.append((p, p.grad, buf if group['momentum'] != 0 else None, group['weight_decay'], group['lr'],
                                group['momentum'], group['dampening'], lars, first))
        self._fused(entries, self.trust_coef, self.eps)
        return loss
