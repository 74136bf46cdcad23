"""Mirror of codes/models/Video_base_model.py:16-251 -- ``VideoBaseModel``: owns ``netG`` (EDVR), the pixel loss, the
optimiser with the reference's parameter groups, and the train / test steps the drivers call
(``feed_data``, ``optimize_parameters``, ``optimize_by_loss``, ``calculate_loss``, ``test``, ``get_current_log``,
``get_current_visuals``, ``load`` / ``save``).

B200-first differences behind the same API:
  * the pixel loss value AND its gradient come from one kernel (``ops.pixel_loss``: l1 / l2 / cb / huber,
    Video_base_model.py:39-50, loss.py:5-30);
  * parameters live in one flat buffer; ``optimizer_G`` is a ``FlatOptimizer`` (one launch per step, same
    ``param_groups`` interface, the reference's effective ``ft_tsa_only`` / ``small_offset_lr`` groupings: ``param_group_spec``);
  * ``log_dict['l_pix']`` is a device scalar, converted to float when the log is read (the reference's ``.item()`` per
    step, :178,:189,:194, is a host sync on the critical path);
  * one process drives one GPU: the DataParallel / DDP wrappers are a ``.module`` pass-through (base_model.py here).
"""
import logging
import os
from collections import OrderedDict

import torch

from . import lr_scheduler, networks
from .. import ops
from ..optim import FlatOptimizer
from .base_model import BaseModel, DataParallel, DistributedDataParallel, NetWrapperMixin

logger = logging.getLogger('base')

_HUBER_DELTA = 1e-2   # loss.py:8


class _PixelCriterion(object):
    """Callable with nn.Module-loss call shape: ``cri_pix(fake, real) -> device scalar`` (gradient flows to ``fake``)."""

    def __init__(self, kind):
        self.kind = kind

    def __call__(self, fake, real):
        return ops.pixel_loss(fake, real.detach(), self.kind, 1.0, _HUBER_DELTA if self.kind == 'huber' else 1e-6)

    def to(self, device):
        return self


class _LogDict(OrderedDict):
    """Device scalars in, floats out: ``log['l_pix']`` materialises (and caches) the value on first read."""

    def __getitem__(self, k):
        v = OrderedDict.__getitem__(self, k)
        if torch.is_tensor(v):
            v = float(v)
            OrderedDict.__setitem__(self, k, v)
        return v

    def items(self):
        return [(k, self[k]) for k in self.keys()]


def param_group_spec(names, train_opt):
    """Optimiser groups ``[(lr, [parameter names]), ...]`` exactly as the reference ends up with them
    (Video_base_model.py:57-130), which is not what its options suggest: the two groups built for ``ft_tsa_only`` (:57-78)
    are overwritten by the ``freeze_front`` / ``small_offset_lr`` / else chain that follows (:79-130).  So

      * ``small_offset_lr``: [others at lr_G, {pcd_align, fea_L*, feature_extraction, conv_first} at 0.1 * lr_G], with or
        without ``ft_tsa_only``;
      * otherwise ONE group with every parameter -- also for ``ft_tsa_only`` alone, where ``set_params_lr_zero`` (:161-163)
        then zeroes the rate of everything for the first ``ft_tsa_only`` iterations (and the chained schedule keeps it at
        zero until a restart: models/lr_scheduler.py here).

    Checked against the unmodified reference class (tests/golden/wrapper_train.json)."""
    lr = train_opt['lr_G']
    if train_opt['freeze_front']:
        raise NotImplementedError('freeze_front targets DUF layer names (Video_base_model.py:79-97); EDVR only here')
    if train_opt['small_offset_lr']:
        slow = lambda k: any(t in k for t in ('pcd_align', 'fea_L', 'feature_extraction', 'conv_first'))
        return [(lr, [k for k in names if not slow(k)]), (lr * 0.1, [k for k in names if slow(k)])]   # "normal params first"
    return [(lr, list(names))]


class VideoBaseModel(NetWrapperMixin, BaseModel):
    def __init__(self, opt):
        super(VideoBaseModel, self).__init__(opt)
        self.rank = torch.distributed.get_rank() if opt['dist'] else -1
        train_opt = opt['train']

        self.netG = networks.define_G(opt).to(self.device)
        self.netG = DistributedDataParallel(self.netG) if opt['dist'] else DataParallel(self.netG)
        self.load()
        self.log_dict = _LogDict()

        loss_type = train_opt['pixel_criterion']
        if loss_type not in ('l1', 'l2', 'cb', 'huber'):
            raise NotImplementedError('Loss type [{:s}] is not recognized.'.format(str(loss_type)))
        self.cri_pix = _PixelCriterion(loss_type)
        self.l_pix_w = train_opt['pixel_weight']

        if self.is_train:
            self.netG.train()
            wd_G = train_opt['weight_decay_G'] if train_opt['weight_decay_G'] else 0
            named = [(k, v) for k, v in self.netG.named_parameters()]
            for k, v in named:
                if not v.requires_grad and self.rank <= 0:
                    logger.warning('Params [{:s}] will not optimize.'.format(k))
            named = [(k, v) for k, v in named if v.requires_grad]
            spec = param_group_spec([k for k, _ in named], train_opt)
            by_name = dict(named)
            if len(spec) == 1:
                optim_params = [by_name[k] for k in spec[0][1]]
            else:
                optim_params = [{'params': [by_name[k] for k in names], 'lr': lr} for lr, names in spec]
            kind = 'SGD' if train_opt['optim'] == 'SGD' else 'Adam'
            betas = (train_opt['beta1'] if train_opt['beta1'] is not None else 0.9,
                     train_opt['beta2'] if train_opt['beta2'] is not None else 0.999)
            self.optimizer_G = FlatOptimizer(optim_params, kind=kind, lr=train_opt['lr_G'], weight_decay=wd_G, betas=betas)
            self.optimizers.append(self.optimizer_G)

            if train_opt['lr_scheme'] == 'MultiStepLR':
                for optimizer in self.optimizers:
                    self.schedulers.append(lr_scheduler.MultiStepLR_Restart(
                        optimizer, train_opt['lr_steps'], restarts=train_opt['restarts'], weights=train_opt['restart_weights'],
                        gamma=train_opt['lr_gamma'] if train_opt['lr_gamma'] is not None else 0.1,
                        clear_state=train_opt['clear_state']))
            elif train_opt['lr_scheme'] == 'CosineAnnealingLR_Restart':
                for optimizer in self.optimizers:
                    self.schedulers.append(lr_scheduler.CosineAnnealingLR_Restart(
                        optimizer, train_opt['T_period'], eta_min=train_opt['eta_min'], restarts=train_opt['restarts'],
                        weights=train_opt['restart_weights']))
            elif train_opt['lr_scheme'] is not None:
                raise NotImplementedError('lr_scheme [{}] is not a DynaVSR scheme'.format(train_opt['lr_scheme']))

    # ------------------------------------------------------------------ data
    def feed_data(self, data, need_GT=True):
        self.var_L = data['LQs'].to(self.device, non_blocking=True)
        if need_GT:
            self.real_H = data['GT'].to(self.device, non_blocking=True)

    def set_params_lr_zero(self):
        self.optimizers[0].param_groups[0]['lr'] = 0

    # ------------------------------------------------------------------ steps
    def optimize_parameters(self, step):
        if self.opt['train']['ft_tsa_only'] and step < self.opt['train']['ft_tsa_only']:
            self.set_params_lr_zero()
        self.optimizer_G.zero_grad()
        self.fake_H = self.netG(self.var_L)
        l_pix = self.l_pix_w * self.cri_pix(self.fake_H, self.real_H)
        l_pix.backward()
        self.optimizer_G.step()
        self.log_dict['l_pix'] = l_pix.detach()

    def optimize_by_loss(self, loss):
        if self.opt['train']['ft_tsa_only']:
            self.set_params_lr_zero()
        self.optimizer_G.zero_grad()
        loss.backward()
        self.optimizer_G.step()
        self.log_dict['l_pix'] = loss.detach()

    def calculate_loss(self):
        self.fake_H = self.netG(self.var_L)
        l_pix = self.l_pix_w * self.cri_pix(self.fake_H, self.real_H)
        self.log_dict['l_pix'] = l_pix.detach()
        return l_pix

    def test(self):
        self.netG.eval()
        with torch.no_grad():
            self.fake_H = self.netG(self.var_L)
        self.netG.train()

    # ------------------------------------------------------------------ visuals (log: NetWrapperMixin)
    def get_current_visuals(self, need_GT=True):
        out = OrderedDict(LQ=self._cpu_frame(self.var_L), rlt=self._cpu_frame(self.fake_H))
        if need_GT:
            out['GT'] = self._cpu_frame(self.real_H)
        return out

    def print_network(self):
        if self.rank <= 0:
            self._log_structure(logger, self.netG, 'G')

    # ------------------------------------------------------------------ checkpoints
    def load(self, verbose=True):
        load_path_G = self.opt['path']['pretrain_model_G']
        if load_path_G is not None:
            if verbose:
                logger.info('Loading model for G [{:s}] ...'.format(load_path_G))
            self.load_network(load_path_G, self.netG, self.opt['path']['strict_load'])

    def load_for_test(self):
        self.load_network(os.path.join(self.opt['path']['models'], 'latest_G.pth'), self.netG, self.opt['path']['strict_load'])

    def save(self, iter_label):
        self.save_network(self.netG, 'G', iter_label)

    def save_for_test(self):
        self.save_network(self.netG, 'G', 'latest')
