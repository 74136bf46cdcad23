"""CPU-side host logic: option dict, LR schedules (against torch and against traces of the reference classes), model-wrapper plumbing."""
import pytest
import torch


def test_nonedict():
    from dynavsr_b200.options import dict_to_nonedict
    o = dict_to_nonedict({'a': {'b': 1}, 'l': [{'c': 2}]})
    assert o['missing'] is None and o['a']['nope'] is None and o['a']['b'] == 1 and o['l'][0]['zz'] is None


class _Opt(object):
    def __init__(self, lrs):
        self.param_groups = [{'lr': lr} for lr in lrs]


def test_multistep_restart_matches_torch_multistep():
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



# ---------------------------------------------------------------------------------------------------------------
# BaseModel plumbing (codes/models/base_model.py:8-121): file names, key prefixes, warm-up ramp, state resume.
def _toy_model(tmp_path):
    import torch.nn as nn
    from dynavsr_b200.models.base_model import BaseModel, DataParallel, NetWrapperMixin
    from dynavsr_b200.models.lr_scheduler import MultiStepLR_Restart

    class Toy(NetWrapperMixin, BaseModel):
        pass

    for sub in ('models', 'training_state'):
        (tmp_path / sub).mkdir(exist_ok=True)
    opt = {'gpu_ids': [0], 'is_train': True,
           'path': {'models': str(tmp_path / 'models'), 'training_state': str(tmp_path / 'training_state')}}
    m = Toy(opt)
    m.net = DataParallel(nn.Sequential(nn.Linear(3, 4), nn.Linear(4, 2)))
    o = torch.optim.SGD(m.net.parameters(), lr=0.1, momentum=0.9)
    m.optimizers = [o]
    m.schedulers = [MultiStepLR_Restart(o, [2, 4], gamma=0.5)]
    m.log_dict = {'l_pix': 1.0}
    return m


def test_base_model_requires_gpu_ids():
    from dynavsr_b200.models.base_model import BaseModel
    with pytest.raises(NotImplementedError):
        BaseModel({'gpu_ids': None, 'is_train': False})


def test_base_model_warmup_and_schedule(tmp_path):
    m = _toy_model(tmp_path)
    lrs = []
    for it in range(1, 7):
        m.update_learning_rate(it, warmup_iter=3)
        lrs.append(m.get_current_learning_rate()[0])
    # iterations 1, 2: the warm-up writes initial_lr * it / 3 over whatever the schedule computed (milestone 2 included); from
    # iteration 3 on the CHAINED schedule carries the last warm-up value (0.2 / 3) forward and halves it at milestone 4 --
    # the reference's behaviour (torch chained schedulers; trace 'multistep_warmup' in tests/golden/lr_schedules.json)
    assert lrs == pytest.approx([0.1 / 3, 0.2 / 3, 0.2 / 3, 0.1 / 3, 0.1 / 3, 0.1 / 3])
    assert m.get_current_log() is m.log_dict


def test_base_model_network_files_and_prefix_stripping(tmp_path):
    import torch.nn as nn
    m = _toy_model(tmp_path)
    m.save_network(m.net, 'G', 1234)
    path = tmp_path / 'models' / '1234_G.pth'
    assert path.exists()
    saved = torch.load(str(path))
    assert list(saved) == ['0.weight', '0.bias', '1.weight', '1.bias']          # wrapper prefix never reaches the file
    assert all(v.device.type == 'cpu' for v in saved.values())
    # a checkpoint written from a torch DataParallel model carries 'module.' keys: they load into a bare network
    prefixed = tmp_path / 'models' / 'dp.pth'
    torch.save({'module.' + k: v + 1 for k, v in saved.items()}, str(prefixed))
    fresh = nn.Sequential(nn.Linear(3, 4), nn.Linear(4, 2))
    storage = fresh[0].weight.data_ptr()
    m.load_network(str(prefixed), fresh)
    assert fresh[0].weight.data_ptr() == storage                                 # copied INTO the existing storage
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, saved[k] + 1)
    with pytest.raises(RuntimeError):
        m.load_network(str(prefixed), nn.Sequential(nn.Linear(3, 4)), strict=True)
    text, count = m.get_network_description(m.net)
    assert count == 3 * 4 + 4 + 4 * 2 + 2 and 'Linear' in text


def test_base_model_training_state_round_trip(tmp_path):
    m = _toy_model(tmp_path)
    loss = m.net(torch.ones(1, 3)).sum()
    loss.backward()
    m.optimizers[0].step()
    for it in range(1, 4):
        m.update_learning_rate(it)
    m.save_training_state(7, 300)
    m.save_training_state(7, 300, model_type='E')
    d = tmp_path / 'training_state'
    assert (d / '300.state').exists() and (d / '300_E.state').exists()
    state = torch.load(str(d / '300.state'))
    assert state['epoch'] == 7 and state['iter'] == 300
    m2 = _toy_model(tmp_path)
    m2.resume_training(state)
    assert m2.get_current_learning_rate() == m.get_current_learning_rate()
    assert m2.schedulers[0].last_epoch == m.schedulers[0].last_epoch
    buf = lambda mm: [s['momentum_buffer'] for s in mm.optimizers[0].state_dict()['state'].values()]
    assert all(torch.equal(a, b) for a, b in zip(buf(m), buf(m2)))
    m2.optimizers.append(m2.optimizers[0])
    with pytest.raises(AssertionError):
        m2.resume_training(state)


def test_net_wrapper_structure_log(tmp_path):
    import logging
    m = _toy_model(tmp_path)
    records = []

    class H(logging.Handler):
        def emit(self, r):
            records.append(r.getMessage())

    lg = logging.getLogger('test_structure')
    lg.setLevel(logging.INFO)
    lg.addHandler(H())
    m._log_structure(lg, m.net, 'G')
    assert 'DataParallel - Sequential' in records[0] and 'parameters: 26' in records[0]


def test_create_model_names_and_order(monkeypatch):
    """codes/models/__init__.py:5-37: 'a+b' -> list in that order, single name -> the wrapper itself, unknown names raise."""
    import dynavsr_b200.models as M
    from dynavsr_b200.models import LRestimator_model, Video_base_model

    class G(object):
        def __init__(self, opt):
            self.opt = opt

    class E(G):
        pass

    monkeypatch.setattr(Video_base_model, 'VideoBaseModel', G)
    monkeypatch.setattr(LRestimator_model, 'LRimgestimator_Model', E)
    opt = {'model': 'video_base+lrimgestimator'}
    both = M.create_model(opt)
    assert [type(m) for m in both] == [G, E] and both[0].opt is opt
    assert [type(m) for m in M.create_model({'model': 'lrimgestimator+video_base'})] == [E, G]
    assert type(M.create_model({'model': 'video_base'})) is G
    for name, text in (('srgan', 'outside the DynaVSR hot path'), ('nope', 'not recognized')):
        with pytest.raises(NotImplementedError, match=text):
            M.create_model({'model': name})


@pytest.mark.parametrize('case', ['cosine_restarts', 'cosine_two_groups', 'multistep_restarts', 'multistep_warmup',
                                  'multistep_zero_group0', 'cosine_warmup'])
def test_lr_schedules_match_reference_traces(case, tmp_path):
    """tests/golden/lr_schedules.json: per-iteration learning rates of the UNMODIFIED reference scheduler classes
    (oracle/make_golden_lr.py), incl. weighted restarts, two parameter groups, running past the last cosine period, and --
    driven through ``BaseModel.update_learning_rate`` -- rates written from outside (warm-up, a group frozen at zero), which
    the reference's chained schedulers carry forward."""
    import json
    import os
    from util import GOLD
    from dynavsr_b200.models import lr_scheduler as S
    c = json.load(open(os.path.join(GOLD, 'lr_schedules.json')))[case]

    def build():
        params = [torch.nn.Parameter(torch.zeros(1)) for _ in c['lrs']]
        opt = torch.optim.SGD([{'params': [p], 'lr': lr} for p, lr in zip(params, c['lrs'])], lr=c['lrs'][0])
        cls = S.CosineAnnealingLR_Restart if c['kind'] == 'cosine' else S.MultiStepLR_Restart
        return opt, cls(opt, **c['kw'])

    opt, sch = build()
    floor = 1e-12 * max(c['lrs'])              # the reference's recursion reaches the troughs with ~1e-20 of rounding residue
    outside = c.get('warmup_iter') is not None or c.get('zero_group0_before') is not None
    if outside:
        m = _toy_model(tmp_path)
        m.optimizers, m.schedulers = [opt], [sch]
    for t, want in enumerate(c['trace']):
        if outside:
            m.update_learning_rate(t + 1, warmup_iter=c.get('warmup_iter') or -1)
            if t + 1 < (c.get('zero_group0_before') or 0):
                opt.param_groups[0]['lr'] = 0
        else:
            sch.step()
        got = [g['lr'] for g in opt.param_groups]
        assert got == pytest.approx(want, rel=1e-9, abs=floor), (t, got, want)
    if not outside:
        # resuming: optimiser groups + scheduler state restored -> the trace continues where it was
        opt2, sch2 = build()
        for _ in range(9):
            sch2.step()
        state_o, state_s = opt2.state_dict(), sch2.state_dict()
        opt3, sch3 = build()
        opt3.load_state_dict(state_o)
        sch3.load_state_dict(state_s)
        for t in range(9, 14):
            sch3.step()
            assert [g['lr'] for g in opt3.param_groups] == pytest.approx(c['trace'][t], rel=1e-9, abs=floor)


def test_param_groups_match_reference_wrapper():
    """tests/golden/wrapper_train.json: optimiser groups (learning rate + parameter names, in order) the UNMODIFIED reference
    VideoBaseModel builds for the ft_tsa_only / small_offset_lr options (oracle/make_golden_wrapper.py)."""
    import json
    import os
    from util import GOLD
    from dynavsr_b200.models.Video_base_model import param_group_spec
    from dynavsr_b200.options import dict_to_nonedict
    from oracle import params as P
    g = json.load(open(os.path.join(GOLD, 'wrapper_train.json')))
    names = ['module.' + k for k in P.edvr_param_shapes(scale=4, **g['net'])]        # the wrapper sees DataParallel names
    assert len(g['cases']) == 5
    for case, c in g['cases'].items():
        spec = param_group_spec(names, dict_to_nonedict(c['train']))
        assert len(spec) == len(c['groups']), case
        for (lr, got), want in zip(spec, c['groups']):
            assert lr == pytest.approx(want['lr']) and [k[len('module.'):] for k in got] == want['names'], case
    # the reference's quirk, spelled out: ft_tsa_only alone does NOT split the parameters
    assert len(g['cases']['ft_tsa_only3_sgd_l1']['groups']) == 1
    assert len(g['cases']['ft_tsa_and_small_offset']['groups']) == 2
    with pytest.raises(NotImplementedError):
        param_group_spec(names, dict_to_nonedict({'lr_G': 1e-4, 'freeze_front': True}))


def test_wrappers_construct_with_reference_option_dicts(monkeypatch):
    """Factory -> wrappers -> optimiser groups -> schedules with the device redirected to the CPU (construction is host logic;
    the steps themselves need the CUDA library and are covered by tests/test_models_gpu.py)."""
    import dynavsr_b200.models.base_model as BM
    from dynavsr_b200.models import create_model
    from dynavsr_b200.options import dict_to_nonedict
    real_device = torch.device

    class CpuTorch(object):
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def device(*a, **k):
            return real_device('cpu')

    monkeypatch.setattr(BM, 'torch', CpuTorch())

    def opt(model, is_train=True, **train):
        t = dict(pixel_criterion='cb', pixel_weight=1.0, optim='Adam', lr_G=1e-4, beta1=0.9, beta2=0.99, lr_scheme='MultiStepLR',
                 lr_steps=[2, 4], lr_gamma=0.5, loss_ftn='l1', lr_C=1e-4)
        t.update(train)
        return dict_to_nonedict({
            'model': model, 'scale': 4, 'gpu_ids': [0], 'dist': False, 'is_train': is_train,
            'network_G': {'which_model_G': 'EDVR', 'nf': 16, 'nframes': 3, 'groups': 2, 'front_RBs': 1, 'back_RBs': 1,
                          'predeblur': False, 'HR_in': False, 'w_TSA': True},
            'network_E': {'which_model_E': 'MFDN', 'mode': 'video', 'nf': 16, 'in_nc': 3},
            'path': {'strict_load': True}, 'train': t})

    g, e = create_model(opt('video_base+lrimgestimator', is_train=False))              # the test-time drivers' call
    assert type(g).__name__ == 'VideoBaseModel' and type(e).__name__ == 'LRimgestimator_Model' and not g.optimizers
    g, e = create_model(opt('video_base+lrimgestimator', ft_tsa_only=5))
    assert [len(grp['params']) for grp in g.optimizer_G.param_groups] == [len(list(g.netG.parameters()))]   # ONE group (quirk)
    assert len(e.optimizer_E.param_groups) == 1 and len(e.optimizer_E.param_groups[0]['params']) == 14
    g.update_learning_rate(1)
    g.update_learning_rate(2)
    assert g.get_current_learning_rate() == [pytest.approx(5e-5)]                       # milestone at 2
    g.set_params_lr_zero()
    g.update_learning_rate(3)
    g.update_learning_rate(4)
    assert g.get_current_learning_rate() == [0]                                         # chained schedule: stays frozen
    c = create_model(opt('video_base', small_offset_lr=True, lr_scheme='CosineAnnealingLR_Restart', T_period=[10, 10],
                         restarts=[10], restart_weights=[1], eta_min=1e-7))
    lrs = [grp['lr'] for grp in c.optimizer_G.param_groups]
    assert lrs == [pytest.approx(1e-4), pytest.approx(1e-5)] and type(c.schedulers[0]).__name__ == 'CosineAnnealingLR_Restart'
    for it in range(1, 11):
        c.update_learning_rate(it)
    assert c.get_current_learning_rate() == [pytest.approx(1e-7), pytest.approx(1e-7)]   # the trough after one period
    with pytest.raises(NotImplementedError):
        create_model(opt('video_base', lr_scheme='StepLR'))
    with pytest.raises(NotImplementedError):
        create_model(opt('lrimgestimator', lr_scheme='CosineAnnealingLR_Restart'))       # the estimator only knows MultiStepLR


def test_conv_precision_context(monkeypatch):
    """ops.conv_precision: temporary operand precision of the resident-weight conv: nests, restores (also on the error path), is a
    no-op for None and under the exact-fp32 backend.  The precision is stamped into each descriptor (dvsr_policy.precision)
    together with the launch policy of the scope that owns the weight -- there is no library-global state to set."""
    from dynavsr_b200 import _lib, ops
    monkeypatch.setitem(ops._backend, 'tc', True)
    monkeypatch.setitem(ops._backend, 'precision', 'bf16x3')
    assert ops._new_desc().policy.precision == _lib.PREC_BF16X3 == 0            # zero-initialised descriptor = default
    with ops.conv_precision('bf16'):
        assert ops._backend['precision'] == 'bf16' and ops._new_desc().policy.precision == _lib.PREC_BF16
        with ops.conv_precision(None):
            assert ops._backend['precision'] == 'bf16'
        with ops.conv_precision('tf32'):
            assert ops._new_desc().policy.precision == _lib.PREC_TF32
        assert ops._backend['precision'] == 'bf16'
    assert ops._backend['precision'] == 'bf16x3'
    with pytest.raises(RuntimeError):
        with ops.conv_precision('bf16'):
            raise RuntimeError('launch failed')
    assert ops._backend['precision'] == 'bf16x3'                       # restored on the error path too
    monkeypatch.setitem(ops._backend, 'tc', False)
    with ops.conv_precision('bf16'):
        assert ops._backend['precision'] == 'bf16x3'
    with pytest.raises(AssertionError):
        ops.set_conv_backend(True, 'fp8')
    # launch policy: per scope, picked up from the weight's owner, independent of the current scope
    a, b = ops.new_scope(ops.LaunchPolicy(37, 2, 24)), ops.new_scope(ops.LaunchPolicy(cta_budget=74, mdcn_staged=2))
    w = torch.zeros(1)
    w._dvsr_scope = a
    with ops.scope(b):
        d_cur, d_own = ops._new_desc(), ops._new_desc(w)
    assert (d_cur.policy.cta_budget, d_cur.policy.min_tiles, d_cur.policy.min_chunks, d_cur.policy.mdcn_staged) == (74, 0, 0, 2)
    assert (d_own.policy.cta_budget, d_own.policy.min_tiles, d_own.policy.min_chunks, d_own.policy.mdcn_staged) == (37, 2, 24, 0)
    p0 = ops._new_desc().policy
    assert (p0.cta_budget, p0.min_tiles, p0.min_chunks, p0.mdcn_staged) == (0, 0, 0, 0)


def test_flat_params_single_owner_and_release():
    """A parameter lives in exactly ONE flat buffer: re-homing it into a second FlatParams must fail loudly instead of
    orphaning the first owner's buffers; release() hands the parameters back, after which the old owner refuses to be used."""
    from dynavsr_b200.adapt import FlatParams
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Conv2d(4, 2, 1))
    want = [p.detach().clone() for p in net.parameters()]
    a = FlatParams([net])
    assert all(p._dvsr_flat() is a and p.data.data_ptr() >= a.flat.data_ptr() for p in net.parameters())
    with pytest.raises(RuntimeError, match='already belongs'):
        FlatParams([net])
    a.release()
    assert all(not hasattr(p, '_dvsr_grad') and not hasattr(p, '_dvsr_flat') for p in net.parameters())
    assert all(torch.equal(p.detach(), w) for p, w in zip(net.parameters(), want))
    with pytest.raises(RuntimeError, match='released'):
        a.restore()
    b = FlatParams([net])                                        # adoptable again
    assert all(p._dvsr_flat() is b for p in net.parameters())
    # gradients that arrive through p.grad are folded into the flat gradient, not ignored
    p0 = next(net.parameters())
    p0.grad = torch.ones_like(p0)
    b.fold_autograd_grads()
    assert p0.grad is None and float(b.grad[:p0.numel()].sum()) == p0.numel()


def test_flat_optimizer_state_round_trip_and_validation():
    """FlatOptimizer.state_dict / load_state_dict: resume restores step, group settings and the flat moments; a state
    of another parameter set, of another optimiser kind, or in torch.optim's per-tensor format is refused (the update kernel would
    index the moments out of bounds / silently drop them)."""
    from dynavsr_b200.optim import FlatOptimizer
    mk = lambda co: torch.nn.Sequential(torch.nn.Conv2d(3, co, 3), torch.nn.Conv2d(co, 2, 1))
    net = mk(4)
    opt = FlatOptimizer([{'params': list(net[0].parameters())}, {'params': list(net[1].parameters()), 'lr': 5e-4}], kind='Adam', lr=1e-3)
    opt.flat.m = torch.arange(opt.flat.numel, dtype=torch.float32)
    opt.flat.v = torch.arange(opt.flat.numel, dtype=torch.float32) * 2
    opt._step = 7
    opt.param_groups[1]['lr'] = 2.5e-4
    sd = opt.state_dict()
    assert all('params' not in g for g in sd['param_groups'])
    net2 = mk(4)
    opt2 = FlatOptimizer([{'params': list(net2[0].parameters())}, {'params': list(net2[1].parameters()), 'lr': 5e-4}], kind='Adam', lr=1e-3)
    opt2.load_state_dict(sd)
    assert opt2._step == 7 and opt2.param_groups[1]['lr'] == 2.5e-4 and opt2.param_groups[0]['lr'] == 1e-3
    assert torch.equal(opt2.flat.m, opt.flat.m) and torch.equal(opt2.flat.v, opt.flat.v)
    assert all(isinstance(p, torch.nn.Parameter) for g in opt2.param_groups for p in g['params'])      # not clobbered by indices
    net3 = mk(8)                                                                                          # another parameter set
    opt3 = FlatOptimizer([{'params': list(net3[0].parameters())}, {'params': list(net3[1].parameters())}], kind='Adam')
    with pytest.raises(ValueError, match='elements'):
        opt3.load_state_dict(sd)
    assert opt3.flat.m is None
    net4 = mk(4)
    with pytest.raises(ValueError, match='kind'):
        FlatOptimizer([{'params': list(net4[0].parameters())}, {'params': list(net4[1].parameters())}], kind='SGD').load_state_dict(sd)
    net5 = mk(4)
    topt = torch.optim.Adam(net5.parameters())
    with pytest.raises(ValueError, match='torch.optim-format'):
        opt2.load_state_dict(topt.state_dict())
    bad = dict(sd)
    bad.pop('exp_avg_sq')
    with pytest.raises(ValueError, match='together'):
        opt2.load_state_dict(bad)
