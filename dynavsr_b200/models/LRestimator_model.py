"""Mirror of codes/models/LRestimator_model.py:28-171 -- ``LRimgestimator_Model``: owns ``netE`` (MFDN or SFDN), maps an LR clip
``data['LQs']`` [B, T, C, H, W] to its super-LR estimate ``fake_L`` [B, T, C, H/s, W/s] (``forward_without_optim`` with
gradients for the inner adaptation step, ``test`` without), and can pre-train the estimator (``optimize_parameters``)."""
import logging
from collections import OrderedDict

import torch

from . import networks
from .. import ops
from ..optim import FlatOptimizer
from .base_model import BaseModel, DataParallel, DistributedDataParallel, NetWrapperMixin
from .lr_scheduler import MultiStepLR_Restart
from .Video_base_model import _LogDict, _PixelCriterion

logger = logging.getLogger('base')


class LRimgestimator_Model(NetWrapperMixin, BaseModel):
    def name(self):
        return 'Estimator_Model'

    def __init__(self, opt):
        super(LRimgestimator_Model, self).__init__(opt)
        net_opt, self.train_opt = opt['network_E'], opt['train']
        self.rank = torch.distributed.get_rank() if opt['dist'] else -1
        self.scale, self.model_name, self.mode = opt['scale'], net_opt['which_model_E'], net_opt['mode']
        if self.mode not in ('image', 'video'):
            raise NotImplementedError("network_E.mode must be 'image' (SFDN) or 'video' (MFDN), got %r" % (self.mode,))
        train_set = (opt.get('datasets') or {}).get('train') or {}
        for key in ('kernel_size', 'patch_size', 'batch_size'):
            setattr(self, key, train_set.get(key))

        wrap = DistributedDataParallel if opt['dist'] else DataParallel
        self.netE = wrap(networks.define_E(opt).to(self.device))
        self.load()
        kind = self.train_opt['loss_ftn']
        self.MyLoss = _PixelCriterion(kind) if kind in ('l1', 'l2') else None
        self.log_dict = _LogDict()
        if self.is_train:
            self._setup_training()

    def _setup_training(self):
        """Adam over the estimator's flat parameter buffer + restartable multi-step schedule (LRestimator_model.py:61-97)."""
        t = self.train_opt
        if t['lr_scheme'] != 'MultiStepLR':
            raise NotImplementedError('MultiStepLR learning rate scheme is enough.')
        self.netE.train()
        trainable = [p for p in self.netE.parameters() if p.requires_grad]
        self.optimizer_E = FlatOptimizer(trainable, kind='Adam', lr=t['lr_C'], weight_decay=t['weight_decay_R'] or 0)
        self.optimizers = [self.optimizer_E]
        gamma = 0.1 if t['lr_gamma'] is None else t['lr_gamma']
        self.schedulers = [MultiStepLR_Restart(o, t['lr_steps'], gamma=gamma) for o in self.optimizers]

    def feed_data(self, data):
        put = lambda t: t.to(self.device, non_blocking=True)
        self.real_H = put(data['LQs'])
        self.real_L = put(data['SuperLQs']) if 'SuperLQs' in data else None
        self.var_H = self.real_H.transpose(1, 2)        # B C T H W (LRestimator_model.py:103)

    def _forward(self):
        # both estimators run frames-major channels-last: 'image' mode (SFDN, LRestimator_model.py:101-102,110-113) treats the
        # B*T frames independently, 'video' mode (MFDN, :103-104,115) convolves over T -- no transposes either way
        B, T, C, H, W = self.real_H.shape
        net = self.netE.module
        frames = ops.to_nhwc(self.real_H.reshape(B * T, C, H, W))
        out = net.forward_nhwc(frames, B, T)
        return ops.to_nchw(out).view(B, T, C, out.shape[1], out.shape[2])

    def optimize_parameters(self, step=None):
        self.optimizer_E.zero_grad()
        self.fake_L = self._forward()
        LR_loss = self.MyLoss(self.fake_L, self.real_L)
        self.log_dict['l_pix'] = LR_loss.detach()
        LR_loss.backward()
        self.optimizer_E.step()

    def forward_without_optim(self, step=None):
        self.fake_L = self._forward()

    def test(self):
        self.netE.eval()
        with torch.no_grad():
            self.fake_L = self._forward()
        self.netE.train()

    def get_current_visuals(self, need_GT=True):
        mid = self.fake_L.size(1) // 2                       # centre frame of the clip
        pick = lambda t: t.detach()[0, mid].float().cpu()
        out = OrderedDict(LQ=pick(self.real_L), rlt=pick(self.fake_L))
        if need_GT:
            out['GT'] = pick(self.real_H)
        return out

    def print_network(self):
        self._log_structure(logger, self.netE, 'R')

    def load(self):
        ckpt = self.opt['path']['pretrain_model_E']
        if ckpt is None:
            return
        logger.info('Loading pretrained model for E [{:s}] ...'.format(ckpt))
        self.load_network(ckpt, self.netE)

    def save(self, iter_step):
        self.save_network(self.netE, 'E', iter_step)
