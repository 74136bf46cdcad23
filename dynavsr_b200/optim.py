"""Fused optimisers over a flat parameter buffer, with torch.optim's interface.

The reference builds ``torch.optim.{SGD,Adam}`` over 144-264 tensors and steps them tensor by tensor
(Video_base_model.py:128-135, test_dynavsr.py:223-231, train_dynavsr.py:175-178).  Here the parameters are re-homed into
ONE flat fp32 buffer (adapt.FlatParams): ``zero_grad`` is one memset, ``step`` ONE kernel launch (dvsr_update_sgd /
dvsr_update_adam), and the weight-gradient kernels accumulate straight into the flat gradient.  ``param_groups`` keep
torch's shape (at most two groups = two learning rates, which covers every grouping the reference builds), so torch LR
schedulers and ``BaseModel._set_lr`` / ``get_current_learning_rate`` work unchanged.
"""
import torch

from . import ops
from .adapt import FlatParams


class FlatOptimizer(torch.optim.Optimizer):
    def __init__(self, params, kind='Adam', lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, scope=None):
        if kind not in ('SGD', 'Adam'):
            raise NotImplementedError(kind)
        super(FlatOptimizer, self).__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > 2:
            raise NotImplementedError('the fused update kernels take two learning-rate groups, got %d' % len(self.param_groups))
        self.kind = kind
        groups = [[p for p in g['params'] if p.requires_grad] for g in self.param_groups]
        self.flat = FlatParams([g for g in groups if g] or groups[:1], scope=scope)
        self._one_group = len([g for g in groups if g]) < 2
        self._step = 0

    def zero_grad(self, set_to_none=True):
        self.flat.zero_grad()
        for g in self.param_groups:
            for p in g['params']:
                p.grad = None

    def _lrs(self):
        lrs = [g['lr'] for g in self.param_groups if any(p.requires_grad for p in g['params'])]
        return (lrs[0], lrs[0]) if len(lrs) < 2 else (lrs[0], lrs[1])

    @torch.no_grad()
    def step(self, closure=None):
        lr0, lr1 = self._lrs()
        g0 = self.param_groups[0]
        self._step += 1
        with ops.scope(self.flat.scope):
            if self.kind == 'SGD':
                self.flat.sgd_step(lr0, lr1, weight_decay=g0['weight_decay'])
            else:
                self.flat.adam_step(lr0, lr1, betas=g0['betas'], eps=g0['eps'], step=self._step,
                                    weight_decay=g0['weight_decay'])

    def state_dict(self):
        sd = {'kind': self.kind, 'step': self._step,
              'param_groups': [{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups]}
        if self.flat.m is not None:
            sd['exp_avg'], sd['exp_avg_sq'] = self.flat.m.cpu(), self.flat.v.cpu()
        return sd

    def load_state_dict(self, sd):
        """Accepts only what ``state_dict()`` of a FlatOptimizer over the SAME parameter set wrote: the moments are flat
        buffers that the update kernel indexes with this optimiser's element count, so a state of another network
        configuration (EDVR-M vs L, a different ``requires_grad`` set) or a torch-format state (per-tensor ``state`` dict,
        ``param_groups`` with ``params`` index lists) is refused instead of being read out of bounds / half-applied."""
        if 'state' in sd or 'kind' not in sd:
            raise ValueError('FlatOptimizer.load_state_dict: this is a torch.optim-format state (per-tensor moments); the flat '
                             'optimiser resumes only from its own state_dict() -- restart the moments or convert them')
        if sd['kind'] != self.kind:
            raise ValueError('FlatOptimizer.load_state_dict: saved optimiser kind %r, this one is %r' % (sd['kind'], self.kind))
        saved_groups = sd.get('param_groups', [])
        if len(saved_groups) != len(self.param_groups):
            raise ValueError('FlatOptimizer.load_state_dict: %d saved param groups, %d here' % (len(saved_groups), len(self.param_groups)))
        m, v = sd.get('exp_avg'), sd.get('exp_avg_sq')
        if (m is None) != (v is None):
            raise ValueError('FlatOptimizer.load_state_dict: exp_avg and exp_avg_sq must come together')
        if m is not None:
            for name, t in (('exp_avg', m), ('exp_avg_sq', v)):
                if not torch.is_tensor(t) or t.dtype != torch.float32 or t.dim() != 1 or t.numel() != self.flat.numel:
                    raise ValueError('FlatOptimizer.load_state_dict: %s must be a float32 vector of %d elements (this parameter '
                                     'set), got %s' % (name, self.flat.numel, tuple(t.shape) if torch.is_tensor(t) else type(t)))
        self._step = int(sd.get('step', 0))
        for g, saved in zip(self.param_groups, saved_groups):
            g.update({k: val for k, val in saved.items() if k != 'params'})
        if m is not None:
            self.flat.m = m.to(self.flat.flat.device).clone()
            self.flat.v = v.to(self.flat.flat.device).clone()
        else:
            self.flat.m = self.flat.v = None
