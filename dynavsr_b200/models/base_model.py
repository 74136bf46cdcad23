"""Mirror of codes/models/base_model.py:8-121 -- device pick, LR warm-up, ``save_network`` / ``load_network`` (strips
``module.``), ``save_training_state`` / ``resume_training``.  The parallel wrapper is a single-device pass-through: one
process drives one GPU (clips are sharded across processes, dynavsr_b200.dist), so nothing is scattered or all-reduced
inside ``forward`` / ``backward`` -- the reference's DataParallel / DDP wrappers (Video_base_model.py:27-31) only add a
``.module`` indirection on one GPU, which is kept so that ``netG.module`` and ``module.``-prefixed checkpoints work."""
import os
from collections import OrderedDict

import torch
import torch.nn as nn


class DataParallel(nn.Module):
    """``.module`` indirection of nn.DataParallel / DistributedDataParallel without scatter / gather / reducer hooks."""

    def __init__(self, module, device_ids=None):
        super(DataParallel, self).__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


DistributedDataParallel = DataParallel


class BaseModel(object):
    def __init__(self, opt):
        self.opt = opt
        if opt['gpu_ids'] is None:
            raise NotImplementedError('dynavsr_b200 models are CUDA-only (no CPU fallback): set gpu_ids')
        self.device = torch.device('cuda')
        self.is_train = opt['is_train']
        self.schedulers = []
        self.optimizers = []

    def feed_data(self, data):
        pass

    def optimize_parameters(self):
        pass

    def get_current_visuals(self):
        pass

    def get_current_losses(self):
        pass

    def print_network(self):
        pass

    def save(self, label):
        pass

    def load(self):
        pass

    # ---- learning rate (base_model.py:37-66)
    def _set_lr(self, lr_groups_l):
        for optimizer, lr_groups in zip(self.optimizers, lr_groups_l):
            for param_group, lr in zip(optimizer.param_groups, lr_groups):
                param_group['lr'] = lr

    def _get_init_lr(self):
        return [[v['initial_lr'] for v in optimizer.param_groups] for optimizer in self.optimizers]

    def update_learning_rate(self, cur_iter, warmup_iter=-1):
        for scheduler in self.schedulers:
            scheduler.step()
        if cur_iter < warmup_iter:
            self._set_lr([[v / warmup_iter * cur_iter for v in init_lr_g] for init_lr_g in self._get_init_lr()])

    def get_current_learning_rate(self):
        return [param_group['lr'] for param_group in self.optimizers[0].param_groups]

    # ---- networks (base_model.py:68-94)
    @staticmethod
    def _unwrap(network):
        return network.module if isinstance(network, (DataParallel, nn.DataParallel)) else network

    def get_network_description(self, network):
        network = self._unwrap(network)
        return str(network), sum(map(lambda x: x.numel(), network.parameters()))

    def save_network(self, network, network_label, iter_label):
        save_path = os.path.join(self.opt['path']['models'], '{}_{}.pth'.format(iter_label, network_label))
        state_dict = OrderedDict((k, v.detach().cpu().clone()) for k, v in self._unwrap(network).state_dict().items())
        torch.save(state_dict, save_path)

    def load_network(self, load_path, network, strict=True):
        network = self._unwrap(network)
        load_net = torch.load(load_path, map_location='cpu')
        clean = OrderedDict((k[7:] if k.startswith('module.') else k, v) for k, v in load_net.items())
        network.load_state_dict(clean, strict=strict)      # copies INTO the existing (possibly flat-buffer) storage
        from .. import ops
        ops.invalidate_weight_cache(getattr(next(iter(network.parameters())), '_dvsr_scope', None))

    # ---- training state (base_model.py:97-121)
    def save_training_state(self, epoch, iter_step, model_type=None):
        state = {'epoch': epoch, 'iter': iter_step, 'schedulers': [s.state_dict() for s in self.schedulers],
                 'optimizers': [o.state_dict() for o in self.optimizers]}
        name = '{}_{}.state'.format(iter_step, model_type) if model_type is not None else '{}.state'.format(iter_step)
        torch.save(state, os.path.join(self.opt['path']['training_state'], name))

    def resume_training(self, resume_state):
        resume_optimizers = resume_state['optimizers']
        resume_schedulers = resume_state['schedulers']
        assert len(resume_optimizers) == len(self.optimizers), 'Wrong lengths of optimizers'
        assert len(resume_schedulers) == len(self.schedulers), 'Wrong lengths of schedulers'
        for i, o in enumerate(resume_optimizers):
            self.optimizers[i].load_state_dict(o)
        for i, s in enumerate(resume_schedulers):
            self.schedulers[i].load_state_dict(s)
