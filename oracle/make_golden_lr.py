"""oracle/make_golden_lr.py -- TEST INFRASTRUCTURE ONLY; run in the build container.

Learning-rate traces of the UNMODIFIED reference schedulers (codes/models/lr_scheduler.py:8-64) stepped once per iteration,
as base_model.py:51-53 does -> tests/golden/lr_schedules.json.  Cases cover restarts with weights, two parameter groups and
running past the end of the last cosine period (the reference's trough-crossing branch, :56-60), and two traces with
outside writes to the rates, driven through the reference ``BaseModel.update_learning_rate`` (base_model.py:44-65): linear
warm-up, and group 0 set to zero for the first iterations (``set_params_lr_zero`` of ``ft_tsa_only``).

    python -m oracle.make_golden_lr
"""
import json
import os
import sys

import torch

REF = os.environ.get('DVSR_REFERENCE', '/root/reference/codes')
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = {
    'cosine_restarts': dict(kind='cosine', lrs=[4e-4], steps=700,
                            kw=dict(T_period=[50, 100, 150], restarts=[50, 150], weights=[1, 0.5], eta_min=1e-7)),
    'cosine_two_groups': dict(kind='cosine', lrs=[2e-4, 5e-5], steps=260,
                              kw=dict(T_period=[40, 60, 80], restarts=[40, 100], weights=[0.7, 1], eta_min=0)),
    'multistep_restarts': dict(kind='multistep', lrs=[1e-3, 2e-4], steps=120,
                               kw=dict(milestones=[20, 40, 70, 90], restarts=[50], weights=[0.5], gamma=0.5)),
    'multistep_warmup': dict(kind='multistep', lrs=[1e-3, 2e-4], steps=12, warmup_iter=4,
                             kw=dict(milestones=[2, 6, 9], gamma=0.5)),
    'multistep_zero_group0': dict(kind='multistep', lrs=[1e-3, 2e-4], steps=12, zero_group0_before=3,
                                  kw=dict(milestones=[5], restarts=[8], weights=[0.5], gamma=0.5)),
    'cosine_warmup': dict(kind='cosine', lrs=[4e-4], steps=40, warmup_iter=5,
                          kw=dict(T_period=[15, 20], restarts=[15], weights=[1], eta_min=1e-7)),
}


def main():
    sys.path.insert(0, REF)
    import models.lr_scheduler as S
    out = {}
    for name, c in CASES.items():
        params = [torch.nn.Parameter(torch.zeros(1)) for _ in c['lrs']]
        opt = torch.optim.SGD([{'params': [p], 'lr': lr} for p, lr in zip(params, c['lrs'])], lr=c['lrs'][0])
        cls = S.CosineAnnealingLR_Restart if c['kind'] == 'cosine' else S.MultiStepLR_Restart
        sch = cls(opt, **c['kw'])
        trace = []
        if 'warmup_iter' in c or 'zero_group0_before' in c:
            from models.base_model import BaseModel
            bm = BaseModel({'gpu_ids': None, 'is_train': True})
            bm.optimizers, bm.schedulers = [opt], [sch]
            for it in range(1, c['steps'] + 1):
                bm.update_learning_rate(it, warmup_iter=c.get('warmup_iter', -1))
                if it < c.get('zero_group0_before', 0):
                    opt.param_groups[0]['lr'] = 0                      # Video_base_model.py:161-167
                opt.step()
                trace.append([g['lr'] for g in opt.param_groups])
        else:
            for _ in range(c['steps']):
                opt.step()
                sch.step()
                trace.append([g['lr'] for g in opt.param_groups])
        out[name] = dict(kind=c['kind'], lrs=c['lrs'], kw=c['kw'], trace=trace, warmup_iter=c.get('warmup_iter'),
                         zero_group0_before=c.get('zero_group0_before'))
    with open(os.path.join(GOLD, 'lr_schedules.json'), 'w') as f:
        json.dump(out, f)
    print({k: len(v['trace']) for k, v in out.items()})


if __name__ == '__main__':
    main()
