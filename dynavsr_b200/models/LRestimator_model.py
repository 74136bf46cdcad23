"""Mirror of codes/models/LRestimator_model.py:28-171 -- ``LRimgestimator_Model``: owns ``netE`` (MFDN), maps an LR clip
``data['LQs']`` [B, T, C, H, W] to its super-LR estimate ``fake_L`` [B, T, C, H/s, W/s] (``forward_without_optim`` with
gradients for the inner adaptation step, ``test`` without), and can pre-train the estimator (``optimize_parameters``)."""
import logging
from collections import OrderedDict

import torch

from . import networks
from .. import ops
from ..optim import FlatOptimizer
from .base_model import BaseModel, DataParallel, DistributedDataParallel
from .lr_scheduler import MultiStepLR_Restart
from .Video_base_model import _LogDict, _PixelCriterion

logger = logging.getLogger('base')


class LRimgestimator_Model(BaseModel):
    def name(self):
        return 'Estimator_Model'

    def __init__(self, opt):
        super(LRimgestimator_Model, self).__init__(opt)
        self.rank = torch.distributed.get_rank() if opt['dist'] else -1
        train_opt = opt['train']
        self.train_opt = train_opt
        ds = (opt.get('datasets') or {}).get('train')
        self.kernel_size = ds['kernel_size'] if ds else None
        self.patch_size = ds['patch_size'] if ds else None
        self.batch_size = ds['batch_size'] if ds else None
        self.scale = opt['scale']
        self.model_name = opt['network_E']['which_model_E']
        self.mode = opt['network_E']['mode']
        if self.mode == 'image':
            raise NotImplementedError("network_E.mode 'image' (SFDN) is not on the DynaVSR-R hot path (MFDN = 'video')")

        self.netE = networks.define_E(opt).to(self.device)
        self.netE = DistributedDataParallel(self.netE) if opt['dist'] else DataParallel(self.netE)
        self.load()

        if train_opt['loss_ftn'] in ('l1', 'l2'):
            self.MyLoss = _PixelCriterion(train_opt['loss_ftn'])
        else:
            self.MyLoss = None
        self.log_dict = _LogDict()

        if self.is_train:
            self.netE.train()
            wd_R = train_opt['weight_decay_R'] if train_opt['weight_decay_R'] else 0
            optim_params = [v for _, v in self.netE.named_parameters() if v.requires_grad]
            self.optimizer_E = FlatOptimizer(optim_params, kind='Adam', lr=train_opt['lr_C'], weight_decay=wd_R)
            self.optimizers = [self.optimizer_E]
            if train_opt['lr_scheme'] == 'MultiStepLR':
                self.schedulers = [MultiStepLR_Restart(o, train_opt['lr_steps'],
                                                       gamma=train_opt['lr_gamma'] if train_opt['lr_gamma'] is not None else 0.1)
                                   for o in self.optimizers]
            else:
                raise NotImplementedError('MultiStepLR learning rate scheme is enough.')

    def feed_data(self, data):
        self.real_H = data['LQs'].to(self.device, non_blocking=True)
        self.real_L = None if 'SuperLQs' not in data.keys() else data['SuperLQs'].to(self.device, non_blocking=True)
        self.var_H = self.real_H.transpose(1, 2)        # B C T H W (LRestimator_model.py:103)

    def _forward(self):
        # MFDN's native layout is frames-major channels-last: skip the two transposes of the reference round trip
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

    def get_current_log(self):
        return self.log_dict

    def get_current_visuals(self, need_GT=True):
        out_dict = OrderedDict()
        T = self.fake_L.size(1)
        out_dict['LQ'] = self.real_L.detach()[0, T // 2].float().cpu()
        out_dict['rlt'] = self.fake_L.detach()[0, T // 2].float().cpu()
        if need_GT:
            out_dict['GT'] = self.real_H.detach()[0, T // 2].float().cpu()
        return out_dict

    def print_network(self):
        s, n = self.get_network_description(self.netE)
        logger.info('Network R structure: {} - {}, with parameters: {:,d}'.format(
            self.netE.__class__.__name__, self.netE.module.__class__.__name__, n))
        logger.info(s)

    def load(self):
        load_path_E = self.opt['path']['pretrain_model_E']
        if load_path_E is not None:
            logger.info('Loading pretrained model for E [{:s}] ...'.format(load_path_E))
            self.load_network(load_path_E, self.netE)

    def save(self, iter_step):
        self.save_network(self.netE, 'E', iter_step)
