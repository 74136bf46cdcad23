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


def test_flat_params_layout_and_views():
    """adapt.FlatParams is plain tensor bookkeeping until a kernel is launched: parameters become views of ONE flat buffer
    (256-byte aligned slots), gradient slots alias ONE flat gradient, the second group starts at `split`, state_dict values and
    keys survive, restore() brings the working copy back to the snapshot."""
    from dynavsr_b200.adapt import FlatParams
    torch.manual_seed(0)
    a, b = torch.nn.Conv2d(3, 5, 3), torch.nn.Conv2d(5, 2, 1)
    before = {k: v.clone() for k, v in list(a.state_dict().items()) + [('b.' + k, v) for k, v in b.state_dict().items()]}
    fp = FlatParams([a, b])
    sizes = [p.numel() for p in list(a.parameters()) + list(b.parameters())]
    assert fp.offsets == [0, 192, 256, 320] and fp.numel == 384 and fp.split == 256      # 135 -> 192, 5 -> 64, 10 -> 64, 2 -> 64
    for p, o, n in zip(fp.params, fp.offsets, sizes):
        assert p.data.data_ptr() == fp.flat.data_ptr() + 4 * o and p._dvsr_grad.data_ptr() == fp.grad.data_ptr() + 4 * o
        assert p._dvsr_grad.shape == p.shape and p._dvsr_scope is fp.scope
    assert all(torch.equal(a.state_dict()[k], before[k]) for k in a.state_dict())
    assert all(torch.equal(b.state_dict()[k], before['b.' + k]) for k in b.state_dict())
    with torch.no_grad():
        a.weight.add_(1.0)                       # an "adaptation step" on the working copy
    assert float((fp.flat - fp.meta).abs().max()) == 1.0
    fp.restore()
    assert torch.equal(fp.flat, fp.meta) and torch.equal(a.weight, before['weight'])
    fp.grad.fill_(2.0)
    fp.zero_grad()
    assert float(fp.grad.abs().max()) == 0.0
    # explicit optimiser groups (Video_base_model.py:57-126): group 0 first, then group 1
    c = torch.nn.Conv2d(3, 4, 1)
    g = FlatParams([[c.bias], [c.weight]])
    assert g.split == 64 and g.params[0] is c.bias and c.weight.data.data_ptr() == g.flat.data_ptr() + 4 * 64

