"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dvsr_b200.h declares (no compute calls without a GPU), and the Python mirror keeps the
reference's names, signatures, parameter inventory and error behaviour."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'dvsr_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dvsr_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from dynavsr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), 'missing export %s' % n
        assert n in _lib.SIGNATURES, 'no ctypes signature for %s' % n
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib().dvsr_version() >= 100


def test_descriptor_struct_sizes_match_header():
    """ctypes mirrors of dvsr_conv_desc / dvsr_wlayout must have the C layout (compiled with gcc)."""
    import subprocess
    import tempfile
    from dynavsr_b200 import _lib
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "dvsr_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",' \
           'sizeof(dvsr_conv_seg),sizeof(dvsr_conv_desc),sizeof(dvsr_wlayout),offsetof(dvsr_conv_desc,Co),offsetof(dvsr_conv_desc,y),' \
           'offsetof(dvsr_conv_desc,policy),sizeof(dvsr_policy));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, 't.c'), 'w').write(prog)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 't.c'), '-o', os.path.join(d, 't')])
        out = subprocess.check_output([os.path.join(d, 't')]).decode().split()
    assert [int(v) for v in out] == [ctypes.sizeof(_lib.ConvSeg), ctypes.sizeof(_lib.ConvDesc), ctypes.sizeof(_lib.WLayout),
                                     _lib.ConvDesc.Co.offset, _lib.ConvDesc.y.offset, _lib.ConvDesc.policy.offset,
                                     ctypes.sizeof(_lib.Policy)]


def test_missing_library_fails_loudly(monkeypatch):
    from dynavsr_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libdvsr_b200.so')
    with pytest.raises(_lib.DvsrError):
        _lib.lib()


def test_reference_api_surface():
    from oracle import params as P
    from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator, arch_util, dcn
    # dcn/__init__.py:1-7 exports; deform_conv.py:99-100,222-223,259 signatures
    for n in ('ModulatedDeformConv', 'ModulatedDeformConvPack', 'modulated_deform_conv'):
        assert hasattr(dcn, n)
    sig = inspect.signature(dcn.ModulatedDeformConvFunction.forward)
    assert list(sig.parameters)[1:] == ['input', 'offset', 'mask', 'weight', 'bias', 'stride', 'padding', 'dilation',
                                        'groups', 'deformable_groups']
    sig = inspect.signature(EDVR_arch.EDVR.__init__)
    assert list(sig.parameters)[1:] == ['nf', 'nframes', 'groups', 'front_RBs', 'back_RBs', 'center', 'predeblur',
                                        'HR_in', 'w_TSA', 'scale']
    # state_dict keys / shapes == reference inventory (asserted against the reference in oracle/make_golden.py)
    for kw, shapes in ((dict(), P.edvr_param_shapes()), (dict(nf=128, back_RBs=40), P.edvr_param_shapes(nf=128, back_RBs=40)),
                       (dict(scale=2), P.edvr_param_shapes(scale=2))):
        sd = EDVR_arch.EDVR(**kw).state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    for scale in (2, 4):
        sd = LRimg_estimator.DirectKernelEstimatorVideo(64, 3, scale).state_dict()
        shapes = P.mfdn_param_shapes(scale=scale)
        assert list(sd.keys()) == list(shapes.keys()) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    # initialisation parity: zero offset conv (deform_conv.py:270-272), 0.1-scaled residual blocks (arch_util.py:46)
    m = dcn.ModulatedDeformConvPack(64, 64, 3, stride=1, padding=1, dilation=1, deformable_groups=8, extra_offset_mask=True)
    assert float(m.conv_offset_mask.weight.abs().sum()) == 0 and float(m.bias.abs().sum()) == 0
    rb = arch_util.ResidualBlock_noBN(64)
    assert 0.3 * 0.1 * (2 / 576) ** 0.5 < float(rb.conv1.weight.std()) < 3 * 0.1 * (2 / 576) ** 0.5
    from oracle import params as P
    for kw in (dict(predeblur=True), dict(HR_in=True, w_TSA=False), dict(predeblur=True, HR_in=True)):
        v = EDVR_arch.EDVR(front_RBs=1, back_RBs=1, **kw)           # constructor variants of EDVR_arch.py:208-239
        full = dict(predeblur=False, HR_in=False, w_TSA=True)
        full.update(kw)
        want = P.edvr_param_shapes(front_RBs=1, back_RBs=1, **full)
        assert [(k, tuple(t.shape)) for k, t in v.state_dict().items()] == [(k, tuple(sh)) for k, sh in want.items()]
    with pytest.raises(NotImplementedError):
        EDVR_arch.EDVR(scale=3)


def test_cpu_tensors_are_rejected_not_emulated():
    from dynavsr_b200.models.archs import EDVR_arch
    net = EDVR_arch.EDVR(front_RBs=1, back_RBs=1)
    with pytest.raises(NotImplementedError):        # deform_conv.py:109-110 behaviour, no CPU fallback
        net(torch.rand(1, 5, 3, 8, 8))
