"""oracle/mdcn_c.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/mdcn_oracle.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libmdcn_oracle.so')
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def forward(x, offset, mask, weight, bias, stride=1, padding=0, dilation=1, groups=1, dg=1):
    x, offset, mask, weight, bias = map(_f, (x, offset, mask, weight, bias))
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    y = np.empty((B, Co, Ho, Wo), np.float32)
    rc = lib().mdcn_oracle_forward(_p(x), _p(offset), _p(mask), _p(weight), _p(bias), _p(y), B, C, H, W, Co,
                                   kh, kw, stride, stride, padding, padding, dilation, dilation, groups, dg)
    if rc:
        raise RuntimeError('mdcn_oracle_forward rc=%d' % rc)
    return y


def backward(x, offset, mask, weight, gy, stride=1, padding=0, dilation=1, groups=1, dg=1):
    x, offset, mask, weight, gy = map(_f, (x, offset, mask, weight, gy))
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    gx, go, gm, gw = (np.empty_like(t) for t in (x, offset, mask, weight))
    gb = np.empty((Co,), np.float32)
    rc = lib().mdcn_oracle_backward(_p(x), _p(offset), _p(mask), _p(weight), _p(gy), _p(gx), _p(go), _p(gm),
                                    _p(gw), _p(gb), B, C, H, W, Co, kh, kw, stride, stride, padding, padding,
                                    dilation, dilation, groups, dg)
    if rc:
        raise RuntimeError('mdcn_oracle_backward rc=%d' % rc)
    return gx, go, gm, gw, gb
