"""Host-side plumbing shared by the model wrappers -- the counterpart of codes/models/base_model.py:8-121: device pick,
learning-rate stepping with linear warm-up, checkpoint files (``{iter}_{label}.pth`` with CPU tensors, ``module.`` prefixes
stripped on load), optimiser / scheduler state files (``{iter}[_{type}].state``) and their restoration.

One process drives one GPU (clips are sharded across processes, dynavsr_b200.dist), so nothing is scattered or all-reduced
inside ``forward`` / ``backward``: the reference's DataParallel / DistributedDataParallel wrappers (Video_base_model.py:27-31)
only add a ``.module`` indirection on a single GPU.  ``DataParallel`` below keeps exactly that indirection (so ``netG.module``
and ``module.``-prefixed checkpoints keep working) and nothing else.
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn


class DataParallel(nn.Module):
    """``wrapper.module`` + pass-through ``forward``; accepts and ignores ``device_ids`` like the torch wrappers."""

    def __init__(self, module, device_ids=None):
        super(DataParallel, self).__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


DistributedDataParallel = DataParallel


def bare(network):
    """The network inside a (torch or local) parallel wrapper."""
    return network.module if isinstance(network, (DataParallel, nn.DataParallel)) else network


class BaseModel(object):
    def __init__(self, opt):
        if opt['gpu_ids'] is None:
            raise NotImplementedError('dynavsr_b200 models are CUDA-only (no CPU fallback): set gpu_ids')
        self.opt = opt
        self.device = torch.device('cuda')
        self.is_train = opt['is_train']
        self.optimizers, self.schedulers = [], []

    # ------------------------------------------------------------------ learning rate
    def update_learning_rate(self, cur_iter, warmup_iter=-1):
        """One scheduler tick per iteration; during warm-up the rate ramps linearly to each group's ``initial_lr``."""
        for sch in self.schedulers:
            sch.step()
        if cur_iter < warmup_iter:
            ramp = cur_iter / float(warmup_iter)
            for opt in self.optimizers:
                for group in opt.param_groups:
                    group['lr'] = group['initial_lr'] * ramp

    def get_current_learning_rate(self):
        return [group['lr'] for group in self.optimizers[0].param_groups]

    # ------------------------------------------------------------------ networks on disk
    def get_network_description(self, network):
        net = bare(network)
        return str(net), sum(p.numel() for p in net.parameters())

    def save_network(self, network, network_label, iter_label):
        target = os.path.join(self.opt['path']['models'], '%s_%s.pth' % (iter_label, network_label))
        torch.save(OrderedDict((k, v.detach().cpu().clone()) for k, v in bare(network).state_dict().items()), target)

    def load_network(self, load_path, network, strict=True):
        net = bare(network)
        weights = OrderedDict()
        for key, value in torch.load(load_path, map_location='cpu').items():
            weights[key[len('module.'):] if key.startswith('module.') else key] = value
        net.load_state_dict(weights, strict=strict)        # copies INTO the existing (possibly flat-buffer) storage
        from .. import ops
        ops.invalidate_weight_cache(getattr(next(iter(net.parameters())), '_dvsr_scope', None))

    # ------------------------------------------------------------------ optimiser / scheduler state on disk
    def save_training_state(self, epoch, iter_step, model_type=None):
        name = '%s.state' % iter_step if model_type is None else '%s_%s.state' % (iter_step, model_type)
        torch.save({'epoch': epoch, 'iter': iter_step,
                    'optimizers': [o.state_dict() for o in self.optimizers],
                    'schedulers': [s.state_dict() for s in self.schedulers]},
                   os.path.join(self.opt['path']['training_state'], name))

    def resume_training(self, resume_state):
        saved_o, saved_s = resume_state['optimizers'], resume_state['schedulers']
        assert len(saved_o) == len(self.optimizers), 'Wrong lengths of optimizers'
        assert len(saved_s) == len(self.schedulers), 'Wrong lengths of schedulers'
        for opt, state in zip(self.optimizers, saved_o):
            opt.load_state_dict(state)
        for sch, state in zip(self.schedulers, saved_s):
            sch.load_state_dict(state)


class NetWrapperMixin(object):
    """What VideoBaseModel and LRimgestimator_Model share: the lazy log, structure printing, label-based save."""

    def get_current_log(self):
        return self.log_dict

    def _log_structure(self, logger, network, tag):
        text, count = self.get_network_description(network)
        logger.info('Network {} structure: {} - {}, with parameters: {:,d}'.format(
            tag, network.__class__.__name__, bare(network).__class__.__name__, count))
        logger.info(text)

    @staticmethod
    def _cpu_frame(t):
        return t.detach()[0].float().cpu()
