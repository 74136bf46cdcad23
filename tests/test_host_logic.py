"""CPU-side host logic: option dict, LR schedule with restarts (closed form vs torch's chained scheduler)."""
import pytest
import torch


def test_nonedict():
    from dynavsr_b200.options import dict_to_nonedict
    o = dict_to_nonedict({'a': {'b': 1}, 'l': [{'c': 2}]})
    assert o['missing'] is None and o['a']['nope'] is None and o['a']['b'] == 1 and o['l'][0]['zz'] is None


class _Opt(object):
    def __init__(self, lrs):
        self.param_groups = [{'lr': lr} for lr in lrs]


def test_multistep_restart_closed_form_matches_torch_multistep():
    from dynavsr_b200.models.lr_scheduler import MultiStepLR_Restart
    ours = _Opt([1e-3, 1e-4])
    sch = MultiStepLR_Restart(ours, [3, 5, 9], gamma=0.5)
    p = [torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))]
    topt = torch.optim.SGD([{'params': [p[0]], 'lr': 1e-3}, {'params': [p[1]], 'lr': 1e-4}], lr=1e-3)
    tsch = torch.optim.lr_scheduler.MultiStepLR(topt, [3, 5, 9], gamma=0.5)
    for _ in range(12):
        assert [g['lr'] for g in ours.param_groups] == pytest.approx([g['lr'] for g in topt.param_groups])
        topt.step()
        tsch.step()
        sch.step()


def test_multistep_restart_weights():
    from dynavsr_b200.models.lr_scheduler import MultiStepLR_Restart
    o = _Opt([1.0])
    sch = MultiStepLR_Restart(o, [2, 7], restarts=[4], weights=[0.5], gamma=0.1)
    seen = []
    for _ in range(9):
        seen.append(o.param_groups[0]['lr'])
        sch.step()
    # t:      0    1    2    3    4    5(restart: 4+1)  6    7     8
    assert seen == pytest.approx([1.0, 1.0, 0.1, 0.1, 0.1, 0.5, 0.5, 0.05, 0.05])
    sd = sch.state_dict()
    sch2 = MultiStepLR_Restart(_Opt([1.0]), [2, 7], restarts=[4], weights=[0.5], gamma=0.1)
    sch2.load_state_dict(sd)
    assert sch2.last_epoch == sch.last_epoch
