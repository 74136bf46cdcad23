"""dynavsr_b200.ops -- autograd operators over the C ABI (include/dvsr_b200.h).

All activations are NHWC fp32 CUDA tensors of shape [N, H, W, C].  PyTorch is used for device memory,
streams and autograd bookkeeping only; every arithmetic pass is a hand-written kernel in
``libdvsr_b200.so``.  There is no CPU path: a non-CUDA tensor raises NotImplementedError exactly as the
reference operator does (codes/models/archs/dcn/deform_conv.py:109-110).
"""
import ctypes
from collections import namedtuple

import torch
from torch.autograd import Function

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID_SPLIT, ConvDesc, WLayout, call

__all__ = ['conv3d_rgb', 'PackScope', 'new_scope', 'scope', 'repack_all', 'snapshot_packs', 'restore_packs', 'invalidate_pack_snapshot', 'weights_updated', 'join_async', 'Seg', 'conv', 'conv3d_padded', 'mdcn', 'upsample', 'pool_maxavg', 'pad2d', 'pad3d_replicate',
           'tsa_temporal', 'tsa_combine', 'pixel_loss', 'to_nhwc', 'to_nchw', 'invalidate_weight_cache',
           'ACT_NONE', 'ACT_RELU', 'ACT_LRELU', 'ACT_SIGMOID_SPLIT', 'LaunchPolicy', 'set_conv_backend', 'conv_precision', 'frame_to_u8']


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise NotImplementedError('dynavsr_b200 ops are CUDA-only (no CPU fallback)')
        if t is not None and t.dtype != torch.float32:
            raise TypeError('dynavsr_b200 ops expect float32 tensors, got %s' % t.dtype)


# --------------------------------------------------------------------------------------------------
# input segments
class Seg(object):
    """One K-segment of a convolution input (see dvsr_conv_seg in include/dvsr_b200.h).

    tensor : [Ns, H, W, C] with stride(3) == 1 (dim 0 / pixel strides may be larger than dense).
    Output image n reads source image (n // T) * Tsrc + (t_fixed if t_fixed >= 0 else n % T + dt).
    """
    __slots__ = ('tensor', 'T', 'Tsrc', 'dt', 't_fixed')

    def __init__(self, tensor, T=1, Tsrc=1, dt=0, t_fixed=-1):
        self.tensor, self.T, self.Tsrc, self.dt, self.t_fixed = tensor, T, Tsrc, dt, t_fixed


def _as_seg(s):
    return s if isinstance(s, Seg) else Seg(s)


def _fill_seg(cs, t, T=1, Tsrc=1, dt=0, t_fixed=-1, ptr_offset=0, img_stride=None):
    assert t.dim() == 4 and t.stride(3) == 1, 'segment must be [N,H,W,C] with unit channel stride'
    assert t.stride(1) == t.shape[2] * t.stride(2), 'segment rows must be dense'
    cs.ptr = t.data_ptr() + 4 * ptr_offset
    cs.C = t.shape[3]
    cs.pix_stride = t.stride(2)
    cs.img_stride = t.stride(0) if img_stride is None else img_stride
    cs.T, cs.Tsrc, cs.dt, cs.t_fixed = T, Tsrc, dt, t_fixed


_SegMeta = namedtuple('_SegMeta', 'T Tsrc dt t_fixed')


# --------------------------------------------------------------------------------------------------
# packed-weight cache: weight tensor object (weakly referenced, so an entry dies with its tensor and a recycled
# device address can never alias a stale pack) -> {layout key: (version stamp, packed buffer)}
import weakref


class _WeightCache(object):
    """id(weight) -> (weakref, {layout key: value}); identity based (tensor == is elementwise)."""

    def __init__(self):
        self._d = {}

    def get(self, weight, key):
        ent = self._d.get(id(weight))
        if ent is None or ent[0]() is not weight:
            return None
        return ent[1].get(key)

    def put(self, weight, key, value):
        ent = self._d.get(id(weight))
        if ent is None or ent[0]() is not weight:
            wid = id(weight)
            ent = (weakref.ref(weight, lambda _r, wid=wid, d=self._d: d.pop(wid, None)), {})
            self._d[wid] = ent
        ent[1][key] = value

    def clear(self):
        self._d.clear()


_wcache = _WeightCache()


class LaunchPolicy(object):
    """Launch policy of the persistent kernels for the calls of ONE engine (dvsr_policy in include/dvsr_b200.h; 0 = library
    default).  It travels inside every descriptor: nothing about a launch is process-global state, so engines / pools with
    different policies can share a process."""
    __slots__ = ('cta_budget', 'min_tiles', 'min_chunks', 'mdcn_staged')

    def __init__(self, cta_budget=0, min_tiles=0, min_chunks=0, mdcn_staged=0):
        self.cta_budget, self.min_tiles, self.min_chunks, self.mdcn_staged = cta_budget, min_tiles, min_chunks, mdcn_staged

    def as_dict(self):
        return {k: getattr(self, k) for k in self.__slots__}


class PackScope(object):
    """Bookkeeping of ONE adaptation engine: its registered weight packs + device-side pack table, the epoch that marks
    packs stale after an out-of-band parameter write, and the side stream its weight-gradient kernels run on.
    Several engines (adapt.AdaptationPool: independent frames in flight on separate streams) each own a scope, so one
    engine's single-launch re-pack never touches another engine's buffers.  Weights are bound to a scope with the
    ``_dvsr_scope`` attribute (adapt.FlatParams); untagged weights live in the process-wide default scope."""

    def __init__(self, policy=None):
        self.policy = policy or LaunchPolicy()
        self.registry = {}      # (id(weight), key) -> entry dict; entries die with their weight (weakref callback)
        self.table = {'dev': None, 'n': 0, 'blocks': 0, 'dirty': True}
        self.epoch = 0
        self.side = None
        self.pending = []
        self.arena = None       # snapshot of every pack buffer (the packs of the meta-weights), see snapshot_packs()
        self.arena_valid = False


_default_scope = PackScope()
_current_scope = [_default_scope]


def new_scope(policy=None):
    return PackScope(policy)


class scope(object):
    """``with ops.scope(s):`` -- makes ``s`` the scope repack_all / weights_updated / join_async act on."""

    def __init__(self, s):
        self.s = s

    def __enter__(self):
        self.prev = _current_scope[0]
        _current_scope[0] = self.s
        return self.s

    def __exit__(self, *exc):
        _current_scope[0] = self.prev
        return False


def _scope_of(weight):
    return getattr(weight, '_dvsr_scope', None) or _default_scope


def _new_desc(weight=None):
    """A zeroed dvsr_conv_desc carrying the launch policy of the engine that owns ``weight`` (its ``_dvsr_scope`` tag; the
    current scope for untagged weights) and the operand precision in force (set_conv_backend / conv_precision)."""
    d = ConvDesc()
    pol = (getattr(weight, '_dvsr_scope', None) or _current_scope[0]).policy
    d.policy.cta_budget, d.policy.min_tiles, d.policy.min_chunks, d.policy.mdcn_staged = \
        pol.cta_budget, pol.min_tiles, pol.min_chunks, pol.mdcn_staged
    d.policy.precision = _PRECISION_CODE[_backend['precision']]
    return d


def invalidate_weight_cache(s=None):
    """Call after parameters were modified through raw pointers; packs are refreshed lazily (one launch each)."""
    (s or _current_scope[0]).epoch += 1


def weights_updated(sc=None):
    """Call after a fused parameter update / restore: bumps the epoch and refreshes every registered pack with one
    table-driven launch (replaces ~370 per-layer pack launches per adaptation step)."""
    sc = sc or _current_scope[0]
    sc.epoch += 1
    repack_all(sc)


def _layout(weight, seg_C, temporal):
    wl = WLayout()
    Co = weight.shape[0]
    if temporal:                      # Conv3d [Co, Ci, KT, KH, KW]: one segment per temporal tap
        _, Ci, KT, KH, KW = weight.shape
        wl.co_stride, wl.ci_stride = Ci * KT * KH * KW, KT * KH * KW
        for s in range(KT):
            wl.seg_base[s], wl.seg_C[s] = s * KH * KW, Ci
        wl.nseg, wl.taps = KT, KH * KW
    else:                             # Conv2d [Co, sum(seg_C), KH, KW]: cat segments
        _, Cin, KH, KW = weight.shape
        assert sum(seg_C) == Cin, 'segment channels %s do not add up to weight in-channels %d' % (seg_C, Cin)
        wl.co_stride, wl.ci_stride = Cin * KH * KW, KH * KW
        off = 0
        for s, c in enumerate(seg_C):
            wl.seg_base[s], wl.seg_C[s] = off * KH * KW, c
            off += c
        wl.nseg, wl.taps = len(seg_C), KH * KW
    wl.Co = Co
    return wl


def _weight_stamp(weight):
    return (weight._version, _scope_of(weight).epoch, weight.data_ptr())


def _get_pack(weight, wl, mode, seg=0, seg_hi=0, a=(0, 0, 0, 0), total=0):
    """Packed copy of ``weight`` in the layout selected by ``mode`` (see dvsr_pack_job).  Packs lazily with one
    launch the first time (or after an out-of-band modification of the weight); afterwards the buffer is kept fresh
    by the table-driven single-launch ``repack_all`` that follows every fused parameter update."""
    key = (mode, seg, seg_hi, a[1], a[2], a[3], tuple(wl.seg_C[i] for i in range(wl.nseg)))
    ent = _wcache.get(weight, key)
    stamp = _weight_stamp(weight)
    if ent is not None and ent['stamp'] == stamp:
        return ent['buf']
    if ent is None:
        job = _lib.PackJob()
        ctypes.memmove(ctypes.byref(job.wl), ctypes.byref(wl), ctypes.sizeof(WLayout))
        job.mode, job.seg, job.seg_hi = mode, seg, seg_hi
        job.a0, job.a1, job.a2, job.a3 = a
        job.total = total
        buf = torch.empty(total, device=weight.device, dtype=torch.float32)
        job.wp = buf.data_ptr()
        sc = _scope_of(weight)
        ent = {'buf': buf, 'job': job, 'stamp': None, 'ref': weakref.ref(weight), 'scope': sc}
        _wcache.put(weight, key, ent)
        rkey = (id(weight), key)
        sc.registry[rkey] = ent
        weakref.finalize(weight, _drop_pack, weakref.ref(sc), rkey)
        sc.table['dirty'] = True
    if ent['job'].w != weight.data_ptr():
        ent['job'].w = weight.data_ptr()
        ent['scope'].table['dirty'] = True
    call('dvsr_pack_job_run', ctypes.byref(ent['job']), _stream())
    ent['stamp'] = stamp
    return ent['buf']


def _drop_pack(scope_ref, rkey):
    sc = scope_ref()
    if sc is not None and sc.registry.pop(rkey, None) is not None:
        sc.table['dirty'] = True


def repack_all(sc=None):
    """Re-pack every weight layout registered in scope ``sc`` (default: the current scope) in ONE launch (dvsr_pack_table)
    and mark the packs fresh."""
    sc = sc or _current_scope[0]
    _pack_table = sc.table
    ents = [e for e in sc.registry.values() if e['ref']() is not None]
    if not ents:
        return
    capturing = torch.cuda.is_current_stream_capturing()
    if _pack_table['dirty'] or _pack_table['n'] != len(ents):
        if capturing:
            raise RuntimeError('weight-pack table changed during CUDA-graph capture (warm up the exact step first)')
        arr = (_lib.PackJob * len(ents))()
        blocks = 0
        for i, e in enumerate(ents):
            e['job'].w = e['ref']().data_ptr()
            e['job'].block_start = blocks
            ctypes.memmove(ctypes.byref(arr[i]), ctypes.byref(e['job']), ctypes.sizeof(_lib.PackJob))
            blocks += (e['job'].total + 255) // 256
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        _pack_table.update(dev=host.to(ents[0]['buf'].device), n=len(ents), blocks=blocks, dirty=False, ents=ents)
        sc.arena_valid = False
    call('dvsr_pack_table', _ptr(_pack_table['dev']), _pack_table['n'], _pack_table['blocks'], _stream())
    for e in _pack_table['ents']:
        w = e['ref']()
        if w is not None:
            e['stamp'] = _weight_stamp(w)


def snapshot_packs():
    """Save every registered pack of the current scope (they must be fresh: call right after repack_all) into one arena,
    so that the next frame can restore them with restore_packs() -- a copy -- instead of re-deriving them."""
    sc = _current_scope[0]
    t = sc.table
    if t['dev'] is None or t['dirty']:
        return False
    n = t['blocks'] * 256
    if sc.arena is None or sc.arena.numel() != n:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError('pack arena must be allocated before CUDA-graph capture')
        sc.arena = torch.empty(n, device=t['dev'].device, dtype=torch.float32)
    call('dvsr_pack_table_copy', _ptr(t['dev']), t['n'], t['blocks'], _ptr(sc.arena), 0, _stream())
    sc.arena_valid = True
    return True


def restore_packs():
    """Bring every pack of the current scope back to the snapshot (one copy launch).  Returns False when there is no
    valid snapshot (new layers were registered, or the meta-weights changed): the caller re-packs instead."""
    sc = _current_scope[0]
    t = sc.table
    ents = [e for e in sc.registry.values() if e['ref']() is not None]
    if not sc.arena_valid or t['dirty'] or t['dev'] is None or t['n'] != len(ents):
        sc.arena_valid = False
        return False
    call('dvsr_pack_table_copy', _ptr(t['dev']), t['n'], t['blocks'], _ptr(sc.arena), 1, _stream())
    for e in t['ents']:
        w = e['ref']()
        if w is not None:
            e['stamp'] = _weight_stamp(w)
    return True


def invalidate_pack_snapshot(s=None):
    (s or _current_scope[0]).arena_valid = False


def _packed(weight, wl, mode, seg=0):
    """CUDA-core layouts -- mode 0: forward [K][Co]; mode 1: data gradient of segment `seg` [taps*Co][C_seg].
    tcgen05 layouts (rows of 32 K-values, K-major, padded N) -- mode 2: forward; mode 3: data gradient."""
    if mode == 0:
        return _get_pack(weight, wl, 0, total=sum(wl.seg_C[i] for i in range(wl.nseg)) * wl.taps * wl.Co)
    if mode == 1:
        return _get_pack(weight, wl, 1, seg, total=wl.seg_C[seg] * wl.taps * wl.Co)
    rows = (wl.Co if mode == 2 else wl.seg_C[seg]) + 15
    rows = rows // 16 * 16
    return _get_pack(weight, wl, mode, seg, a=(rows, 0, 0, 0),
                     total=_lib.lib().dvsr_conv_tc_packed_floats(ctypes.byref(wl), mode, seg))


def _packed_parity(weight, wl, seg, KH, KW, a, b):
    """tcgen05 data-gradient weights restricted to the taps (a + 2t, b + 2u): one parity class of a stride-2 conv."""
    KHs, KWs = (KH - a + 1) // 2, (KW - b + 1) // 2
    rows = (wl.seg_C[seg] + 15) // 16 * 16
    return _get_pack(weight, wl, 4, seg, a=(rows, KW, KWs, 2 * a + b), total=KHs * KWs * ((wl.Co + 31) // 32) * rows * 32)


def _packed_tc2(weight, wl, mode, seg_lo, seg_hi):
    """Resident-weight layout of conv_tc2.cu (mode 5 forward over segments [seg_lo, seg_hi), mode 6 data gradient).
    BF16x3 precision uses modes 9 / 10: 16 KiB blocks per (64-channel pair, tap) with the hi and lo parts stacked along N."""
    bf = _backend['precision'] in ('bf16x3', 'bf16')          # 'bf16' reads the same packs (their hi rows only)
    kdiv = 64 if bf else 32
    if mode == 5:
        nblocks = sum(wl.taps * ((wl.seg_C[s] + kdiv - 1) // kdiv) for s in range(seg_lo, seg_hi))
    else:
        nblocks = wl.taps * ((wl.Co + kdiv - 1) // kdiv)
    if bf:
        mode += 4               # 9 / 10
    return _get_pack(weight, wl, mode, seg_lo, seg_hi, a=(nblocks, 0, 0, 0),
                     total=_lib.lib().dvsr_conv_tc2_packed_floats(ctypes.byref(wl), mode, seg_lo, seg_hi))


def _dgrad_stride2_tc(gpre, weight, wl, seg, shape, spec):
    """Data gradient of a stride-2 convolution as (up to) four stride-1 tensor-core convolutions over gy, one per
    parity class of the input pixel: gx[2i + a - pad] = sum_t gy[i - t] * W[a + 2t] (same along x)."""
    N, H, W, C = shape
    # a 1-wide kernel has an empty odd parity class: those input pixels receive no gradient
    gx = (torch.zeros if (spec.KH < 2 or spec.KW < 2) else torch.empty)(shape, device=gpre.device, dtype=torch.float32)
    for a in (0, 1):
        for b in (0, 1):
            KHs, KWs = (spec.KH - a + 1) // 2, (spec.KW - b + 1) // 2
            off_y, off_x = a - spec.pad, b - spec.pad
            Hi, Wi = (H - 1 - off_y) // 2 + 1, (W - 1 - off_x) // 2 + 1
            if KHs == 0 or KWs == 0:
                continue
            d = _new_desc(weight)
            d.N, d.H, d.W, d.Ho, d.Wo = N, spec.Ho, spec.Wo, Hi, Wi
            d.KH, d.KW, d.stride, d.pad, d.dil, d.transposed = KHs, KWs, 1, 0, 1, 1
            d.nseg = 1
            _fill_seg(d.seg[0], gpre)
            d.Co = C
            d.y, d.y_pix_stride = gx.data_ptr(), C
            d.out_step, d.out_off_y, d.out_off_x, d.out_H, d.out_W = 2, off_y, off_x, H, W
            if _lib.PROFILE['on']:
                _lib.PROFILE['tag'] = 'dgrad-s2[%d%d] %dx%dx%d C%d->%d k%d' % (a, b, N, Hi, Wi, gpre.shape[3], C, spec.KH)
            call('dvsr_conv_tc_fprop', ctypes.byref(d), _ptr(_packed_parity(weight, wl, seg, spec.KH, spec.KW, a, b)), _stream())
    return gx


# --------------------------------------------------------------------------------------------------
# convolution
_backend = {'tc': None, 'precision': 'bf16x3'}     # 'tc': None = auto (tcgen05 when the library exports it), True / False = forced


_PRECISION_CODE = {'bf16x3': _lib.PREC_BF16X3, 'tf32': _lib.PREC_TF32, 'bf16': _lib.PREC_BF16}


def _tc():
    """Is the tcgen05 implicit-GEMM path selected?  It is the DEFAULT whenever the library has the entry points; the
    exact-fp32 CUDA-core path is an explicit choice (``set_conv_backend(False)``: parity studies, op-level tests)."""
    if _backend['tc'] is None:
        _backend['tc'] = hasattr(_lib.lib(), 'dvsr_conv_tc2_fprop')
    return _backend['tc']


def set_conv_backend(tensor_cores, precision=None):
    """Select the tcgen05 implicit-GEMM path (True, the default), the exact-fp32 CUDA-core path (False) or the library default
    (None).  ``precision`` of the resident-weight kernel: 'bf16x3' (default; split operands, 3 products, fp32-class accuracy),
    'tf32' (single pass, ~3e-4 per layer) or 'bf16' (the bf16x3 layouts with only the hi.hi product issued: ~3e-3 per layer, a
    third of the tensor-core work).  The choice travels in every descriptor (dvsr_policy.precision), not in library state."""
    _backend['tc'] = None if tensor_cores is None else bool(tensor_cores)
    if precision is not None:
        assert precision in _PRECISION_CODE
        _backend['precision'] = precision


class conv_precision(object):
    """``with ops.conv_precision('bf16'):`` -- operand precision of the resident-weight tensor-core convolution for the
    launches issued (or captured into a CUDA graph) inside the block; restored on exit.  ``None`` and the exact-fp32 backend
    make it a no-op.  The precision is stamped into each descriptor when the launch is issued."""

    def __init__(self, precision):
        self.precision, self.previous = precision, None

    def __enter__(self):
        if self.precision is not None and _tc() and self.precision != _backend['precision']:
            self.previous = _backend['precision']
            set_conv_backend(True, self.precision)
        return self

    def __exit__(self, *exc):
        if self.previous is not None:
            set_conv_backend(_backend['tc'], self.previous)
            self.previous = None
        return False


# Weight gradients are leaves of the backward pass: when they accumulate into the flat gradient buffer nobody reads
# them before the optimiser step, so they run on a side stream, concurrently with the data-gradient chain (the small
# inner-loop layers use 30-120 CTAs each and leave most SMs idle).  join_async() is called before the update.
_async = {'on': True}


def _side_stream(sc=None):
    sc = sc or _current_scope[0]
    if sc.side is None:
        sc.side = torch.cuda.Stream()
    return sc.side


def join_async(sc=None):
    """Make the current stream wait for every weight-gradient kernel issued on the side stream of scope ``sc`` (default: the
    current scope).  FlatParams joins its OWN scope, which is also the scope the kernels were queued under (the weight's
    ``_dvsr_scope`` tag), so a backward pass run outside ``with ops.scope(...)`` is still ordered before the update."""
    sc = sc or _current_scope[0]
    if sc.pending:
        torch.cuda.current_stream().wait_stream(_side_stream(sc))
        sc.pending.clear()


def _run_wgrad(d, gpre, Co, gw, wl, keep=None, force_tc=False, weight=None):
    """gw += A^T gy with the tensor-core kernel when every segment qualifies, else the CUDA-core kernel.
    ``keep`` (tensors the kernel reads) switches on side-stream execution on the stream of the scope that owns ``weight``;
    they stay referenced until that scope's join_async()."""
    if keep is not None and _async['on'] and not _lib.PROFILE['on']:
        sc = getattr(weight, '_dvsr_scope', None) or _current_scope[0]
        side = _side_stream(sc)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _run_wgrad(d, gpre, Co, gw, wl, force_tc=force_tc)
        sc.pending.append(keep)
        return
    L = _lib.lib()
    if _lib.PROFILE['on']:
        _lib.PROFILE['tag'] = 'wgrad %dx%dx%d C%s->%d k%d s%d%s' % (d.N, d.Ho, d.Wo, '+'.join(str(d.seg[i].C) for i in range(d.nseg)),
                                                                 d.Co, d.KH, d.stride, ' dcn' if d.deform else '')
    if (_tc() or force_tc) and all(L.dvsr_conv_wgrad_tc_supported(ctypes.byref(d), s) == 1 for s in range(d.nseg)):
        for s in range(d.nseg):
            call('dvsr_conv_wgrad_tc', ctypes.byref(d), s, _ptr(gpre), Co, _ptr(gw), ctypes.byref(wl), _stream())
    else:
        call('dvsr_conv_wgrad', ctypes.byref(d), _ptr(gpre), Co, _ptr(gw), ctypes.byref(wl), _stream())


class _ConvSpec(object):
    __slots__ = ('KH', 'KW', 'stride', 'pad', 'act', 'slope', 'sig_split', 'shuffle', 'metas', 'temporal',
                 'N', 'H', 'W', 'Ho', 'Wo', 'has_res')


def _copy_desc(d):
    c = ConvDesc()
    ctypes.memmove(ctypes.byref(c), ctypes.byref(d), ctypes.sizeof(ConvDesc))
    return c


def _try_tc2(d, weight, wl, data_grad, segs):
    """Run ``d`` on the persistent resident-weight kernel, K-splitting over segments when the whole weight set does
    not fit in shared memory.  Returns False if the shape is not eligible."""
    L = _lib.lib()
    if L.dvsr_conv_tc2_supported(ctypes.byref(d)) == 1:
        if data_grad:
            if len(segs) != 1:
                return False
            wp = _packed_tc2(weight, wl, 6, segs[0], segs[0] + 1)
        else:
            wp = _packed_tc2(weight, wl, 5, 0, d.nseg)
        call('dvsr_conv_tc2_fprop', ctypes.byref(d), _ptr(wp), None, 0, _stream())
        return True
    if d.nseg < 2 or d.wshare or d.accumulate or d.out_step or d.deform:
        return False
    subs = []
    for i in range(d.nseg):
        sd = _copy_desc(d)
        sd.nseg = 1
        sd.seg[0] = d.seg[i]
        if L.dvsr_conv_tc2_supported(ctypes.byref(sd)) != 1:
            return False
        subs.append(sd)
    # K-split: partial sums of all but the last segment go to a dense scratch tensor, the last launch adds them
    part = torch.empty(d.N * d.Ho * d.Wo * d.Co, device=weight.device, dtype=torch.float32)
    for i, sd in enumerate(subs):
        last = i == len(subs) - 1
        wp = _packed_tc2(weight, wl, 6, segs[i], segs[i] + 1) if data_grad else _packed_tc2(weight, wl, 5, i, i + 1)
        if not last:
            sd.bias, sd.act, sd.res, sd.shuffle = None, ACT_NONE, None, 0
            sd.y, sd.y_pix_stride = part.data_ptr(), d.Co
        call('dvsr_conv_tc2_fprop', ctypes.byref(sd), _ptr(wp), _ptr(part) if i > 0 else None, d.Co, _stream())
    return True


def _use_tc(d):
    return _tc() and _lib.lib().dvsr_conv_tc_supported(ctypes.byref(d)) == 1


def _run_conv(d, weight, wl, data_grad=False, segs=(0,)):
    """Launch descriptor ``d`` with ``weight`` packed for the chosen kernel family.  ``data_grad`` selects
    the mirrored (mode 1 / 3) weight layout; ``segs`` lists the forward segments whose packed blocks are
    concatenated (one per temporal tap of a Conv3d data gradient)."""
    tc = _use_tc(d)
    if _lib.PROFILE['on']:
        _lib.PROFILE['tag'] = '%s %dx%dx%d C%s->%d k%d s%d' % ('dgrad' if data_grad else 'fprop', d.N, d.Ho, d.Wo,
                                                            '+'.join(str(d.seg[i].C) for i in range(d.nseg)), d.Co, d.KH, d.stride)
    if _tc() and _backend.get('tc2', True) and _try_tc2(d, weight, wl, data_grad, segs):
        return
    if data_grad:
        mode = 3 if tc else 1
        bufs = [_packed(weight, wl, mode, sgi) for sgi in segs]
        wp = bufs[0] if len(bufs) == 1 else torch.cat(bufs)
    else:
        wp = _packed(weight, wl, 2 if tc else 0)
    if tc:
        call('dvsr_conv_tc_fprop', ctypes.byref(d), _ptr(wp), _stream())
    elif d.Co <= 4 and d.nseg == 1 and not d.deform and not d.transposed and not d.shuffle \
            and d.seg[0].C % 4 == 0 and d.seg[0].pix_stride % 4 == 0:
        call('dvsr_conv_small_co', ctypes.byref(d), _ptr(wp), _stream())
    else:
        call('dvsr_conv_fprop', ctypes.byref(d), _ptr(wp), _stream())


def _fwd_desc(spec, tensors, weight=None):
    d = _new_desc(weight)
    d.N, d.H, d.W, d.Ho, d.Wo = spec.N, spec.H, spec.W, spec.Ho, spec.Wo
    d.KH, d.KW, d.stride, d.pad, d.dil = spec.KH, spec.KW, spec.stride, spec.pad, 1
    d.nseg = len(tensors)
    for i, (t, m) in enumerate(zip(tensors, spec.metas)):
        _fill_seg(d.seg[i], t, m.T, m.Tsrc, m.dt, m.t_fixed)
    return d


class _ConvFn(Function):
    @staticmethod
    def forward(ctx, spec, weight, bias, res, *tensors):
        _check_cuda(weight, bias, res, *tensors)
        Co = weight.shape[0]
        wl = _layout(weight, [t.shape[3] for t in tensors], spec.temporal)
        d = _fwd_desc(spec, tensors, weight)
        d.Co = Co
        d.bias = bias.data_ptr() if bias is not None else None
        d.act, d.slope, d.sig_split, d.shuffle = spec.act, spec.slope, spec.sig_split, spec.shuffle
        if spec.shuffle:
            y = torch.empty(spec.N, 2 * spec.Ho, 2 * spec.Wo, Co // 4, device=weight.device, dtype=torch.float32)
            d.y_pix_stride = Co // 4
        else:
            y = torch.empty(spec.N, spec.Ho, spec.Wo, Co, device=weight.device, dtype=torch.float32)
            d.y_pix_stride = Co
        if res is not None:
            assert res.is_contiguous() and res.shape == y.shape
            d.res, d.res_pix_stride = res.data_ptr(), res.shape[3]
        d.y = y.data_ptr()
        _run_conv(d, weight, wl)
        ctx.spec, ctx.wl = spec, wl
        ctx.has_bias, ctx.has_res = bias is not None, res is not None
        ctx.wslot, ctx.bslot = getattr(weight, '_dvsr_grad', None), getattr(bias, '_dvsr_grad', None)
        with_act = spec.act != ACT_NONE
        ctx.save_for_backward(weight, y if with_act else None, res if (with_act and res is not None) else None, *tensors)
        return y

    @staticmethod
    def backward(ctx, gy):
        spec, wl = ctx.spec, ctx.wl
        weight, y, res = ctx.saved_tensors[0], ctx.saved_tensors[1], ctx.saved_tensors[2]
        tensors = ctx.saved_tensors[3:]
        gy = gy.contiguous()
        Co = weight.shape[0]
        npix = spec.N * spec.Ho * spec.Wo
        need_w, need_b, need_res = ctx.needs_input_grad[1], ctx.needs_input_grad[2] and ctx.has_bias, \
            ctx.needs_input_grad[3] and ctx.has_res
        # 1. through the activation (and un-PixelShuffle), bias gradient
        # parameters re-homed by adapt.FlatParams accumulate straight into the flat gradient buffer
        gb = (ctx.bslot if ctx.bslot is not None else torch.zeros(Co, device=gy.device, dtype=torch.float32)) \
            if need_b else None
        if spec.act != ACT_NONE or spec.shuffle:
            gpre = torch.empty(spec.N, spec.Ho, spec.Wo, Co, device=gy.device, dtype=torch.float32)
            call('dvsr_act_bwd', _ptr(gy), _ptr(y), _ptr(res), _ptr(gpre), _ptr(gb), npix, Co, spec.act, spec.slope,
                 spec.sig_split, spec.shuffle, spec.Ho, spec.Wo, _stream())
        else:
            gpre = gy
            if need_b:
                call('dvsr_act_bwd', _ptr(gy), None, None, None, _ptr(gb), npix, Co, ACT_NONE, 0.0, 0, 0, spec.Ho,
                     spec.Wo, _stream())
        # 2. weight gradient
        gw = None
        if need_w:
            gw = ctx.wslot if ctx.wslot is not None else torch.zeros_like(weight)
            d = _fwd_desc(spec, tensors, weight)
            d.Co = Co
            _run_wgrad(d, gpre, Co, gw, wl, keep=(gpre, gy, tensors) if ctx.wslot is not None else None, weight=weight)
        if ctx.wslot is not None:
            gw = None
        if ctx.bslot is not None:
            gb = None
        # 3. data gradients, one transposed convolution per input tensor that needs one
        gts = [None] * len(tensors)
        frame = spec.Ho * spec.Wo * Co
        if spec.temporal:
            if ctx.needs_input_grad[4]:
                t, m0 = tensors[0], spec.metas[0]
                KT = len(tensors)
                gx = torch.empty_like(t)
                d = _new_desc(weight)
                d.N, d.H, d.W, d.Ho, d.Wo = t.shape[0], spec.Ho, spec.Wo, spec.H, spec.W
                d.KH, d.KW, d.stride, d.pad, d.dil, d.transposed = spec.KH, spec.KW, spec.stride, spec.pad, 1, 1
                d.nseg = KT
                for kt in range(KT):
                    _fill_seg(d.seg[kt], gpre, T=m0.Tsrc, Tsrc=m0.T, dt=-kt)
                d.Co = t.shape[3]
                d.y, d.y_pix_stride = gx.data_ptr(), t.shape[3]
                _run_conv(d, weight, wl, data_grad=True, segs=tuple(range(KT)))
                gts[0] = gx
        else:
            for i, (t, m) in enumerate(zip(tensors, spec.metas)):
                if not ctx.needs_input_grad[4 + i]:
                    continue
                if spec.stride == 2 and _tc() and m.T == 1 and m.Tsrc == 1 and Co % 4 == 0 and Co >= 16 \
                        and 16 <= t.shape[3] <= 256 and t.shape[3] % 4 == 0:
                    gts[i] = _dgrad_stride2_tc(gpre, weight, wl, i, tuple(t.shape), spec)
                    continue
                gx = torch.empty(t.shape, device=gy.device, dtype=torch.float32)
                d = _new_desc(weight)
                d.N, d.H, d.W, d.Ho, d.Wo = t.shape[0], spec.Ho, spec.Wo, spec.H, spec.W
                d.KH, d.KW, d.stride, d.pad, d.dil, d.transposed = spec.KH, spec.KW, spec.stride, spec.pad, 1, 1
                if m.T == 1 and m.Tsrc == 1:
                    d.nseg = 1
                    _fill_seg(d.seg[0], gpre)
                elif m.Tsrc == 1 and m.t_fixed == 0:
                    # broadcast input (one source image per clip of T output images): sum over the clip
                    assert m.T <= _lib.MAX_SEG, 'broadcast over %d frames exceeds DVSR_MAX_SEG' % m.T
                    d.nseg, d.wshare = m.T, 1
                    for f in range(m.T):
                        _fill_seg(d.seg[f], gpre, ptr_offset=f * frame, img_stride=m.T * frame)
                else:
                    raise NotImplementedError('data gradient for segment mapping %s' % (m,))
                d.Co = t.shape[3]
                d.y, d.y_pix_stride = gx.data_ptr(), t.shape[3]
                _run_conv(d, weight, wl, data_grad=True, segs=(i,))
                gts[i] = gx
        return (None, gw, gb, gy if need_res else None) + tuple(gts)


def conv(srcs, weight, bias=None, stride=1, pad=1, act=ACT_NONE, slope=0.1, sig_split=0, res=None, shuffle=0):
    """NHWC convolution over one or more input segments with a fused epilogue.

    y = act(conv(cat(srcs), weight) + bias) [+ res], optionally stored through PixelShuffle(2).
    Replaces nn.Conv2d (+ torch.cat, activation, residual add, PixelShuffle) call sites of
    EDVR_arch.py / arch_util.py / LRimg_estimator.py.
    """
    segs = [_as_seg(s) for s in (srcs if isinstance(srcs, (list, tuple)) else [srcs])]
    t0, m0 = segs[0].tensor, segs[0]
    spec = _ConvSpec()
    spec.KH, spec.KW = weight.shape[2], weight.shape[3]
    spec.stride, spec.pad, spec.act, spec.slope, spec.sig_split, spec.shuffle = stride, pad, act, slope, sig_split, shuffle
    spec.metas = [_SegMeta(s.T, s.Tsrc, s.dt, s.t_fixed) for s in segs]
    spec.temporal = False
    spec.H, spec.W = t0.shape[1], t0.shape[2]
    spec.N = (t0.shape[0] // m0.Tsrc) * m0.T
    spec.Ho = (spec.H + 2 * pad - spec.KH) // stride + 1
    spec.Wo = (spec.W + 2 * pad - spec.KW) // stride + 1
    for s in segs:
        assert s.tensor.shape[1] == spec.H and s.tensor.shape[2] == spec.W, 'segment spatial sizes differ'
    return _ConvFn.apply(spec, weight, bias, res, *[s.tensor for s in segs])


def conv3d_padded(xpad, weight, bias, T, act=ACT_NONE, slope=0.1):
    """Conv3d(k=3, pad=0) on an explicitly padded clip tensor.

    xpad: [B*(T+2), H+2, W+2, C] (frames of the replication-padded clip); weight: [Co, C, 3, 3, 3].
    Returns [B*T, H, W, Co].  Implemented as three temporal segments of one implicit GEMM
    (LRimg_estimator.py:77,87,102,112).
    """
    KT, KH, KW = weight.shape[2:]
    spec = _ConvSpec()
    spec.KH, spec.KW, spec.stride, spec.pad = KH, KW, 1, 0
    spec.act, spec.slope, spec.sig_split, spec.shuffle = act, slope, 0, 0
    spec.metas = [_SegMeta(T, T + KT - 1, kt, -1) for kt in range(KT)]
    spec.temporal = True
    spec.H, spec.W = xpad.shape[1], xpad.shape[2]
    spec.N = (xpad.shape[0] // (T + KT - 1)) * T
    spec.Ho, spec.Wo = spec.H - KH + 1, spec.W - KW + 1
    # the same tensor feeds every temporal tap; autograd sees it once (data gradient is returned for slot 0)
    return _ConvFn.apply(spec, weight, bias, None, xpad, *([xpad.detach()] * (KT - 1)))


class _Conv3dRgbFn(Function):
    """Conv3d(C <= 4 -> Co, 3x3x3) over a replication-padded clip (MFDN conv0, LRimg_estimator.py:75-77,100-102) as ONE
    tensor-core 3x3 conv: ``dvsr_tcat_pad3`` folds the three temporal taps into the channel axis of a [B*T, H+2, W+2, 12]
    tensor (channel 4*kt + c) and the weight is addressed through an interleaved-channel ``dvsr_wlayout``.  The weight
    gradient runs on the tensor-core kernel over the same tensor; the clip itself gets no gradient (it is input data)."""

    @staticmethod
    def forward(ctx, x, weight, bias, T, act, slope):
        _check_cuda(x, weight, bias)
        x = x.contiguous()
        BT, H, W, C = x.shape
        Co, Ci, KT, KH, KW = weight.shape
        assert Ci == C and C <= 4 and (KT, KH, KW) == (3, 3, 3) and BT % T == 0
        xc = torch.empty(BT, H + 2, W + 2, 12, device=x.device, dtype=torch.float32)
        call('dvsr_tcat_pad3', _ptr(x), _ptr(xc), BT // T, T, H, W, C, _stream())
        wl = WLayout()
        wl.co_stride, wl.ci_stride, wl.ci_hi_stride = Ci * 27, 27, 9
        wl.ci_bits, wl.ci_lo_valid = 2, C
        wl.seg_base[0], wl.seg_C[0] = 0, 12
        wl.nseg, wl.taps, wl.Co = 1, 9, Co
        d = _new_desc(weight)
        d.N, d.H, d.W, d.Ho, d.Wo = BT, H + 2, W + 2, H, W
        d.KH, d.KW, d.stride, d.pad, d.dil = 3, 3, 1, 0, 1
        d.nseg = 1
        _fill_seg(d.seg[0], xc)
        d.Co = Co
        d.bias = bias.data_ptr() if bias is not None else None
        d.act, d.slope = act, slope
        y = torch.empty(BT, H, W, Co, device=x.device, dtype=torch.float32)
        d.y, d.y_pix_stride = y.data_ptr(), Co
        if _lib.PROFILE['on']:
            _lib.PROFILE['tag'] = 'fprop conv3d-rgb %dx%dx%d C12->%d' % (BT, H, W, Co)
        if _lib.lib().dvsr_conv_tc2_supported(ctypes.byref(d)) != 1:
            raise RuntimeError('conv3d_rgb: shape not supported by the resident-weight tensor-core kernel')
        call('dvsr_conv_tc2_fprop', ctypes.byref(d), _ptr(_packed_tc2(weight, wl, 5, 0, 1)), None, 0, _stream())
        ctx.cfg = (act, slope, BT, H, W, Co, bias is not None)
        ctx.wl = wl
        ctx.wslot, ctx.bslot = getattr(weight, '_dvsr_grad', None), getattr(bias, '_dvsr_grad', None)
        ctx.save_for_backward(xc, weight, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('conv3d_rgb: no gradient w.r.t. the input clip (use conv3d_padded)')
        xc, weight, y = ctx.saved_tensors
        act, slope, BT, H, W, Co, has_bias = ctx.cfg
        gy = gy.contiguous()
        need_b = has_bias and ctx.needs_input_grad[2]
        gb = (ctx.bslot if ctx.bslot is not None else torch.zeros(Co, device=gy.device, dtype=torch.float32)) if need_b else None
        if act != ACT_NONE:
            gpre = torch.empty_like(gy)
            call('dvsr_act_bwd', _ptr(gy), _ptr(y), None, _ptr(gpre), _ptr(gb), BT * H * W, Co, act, slope, 0, 0, H, W, _stream())
        else:
            gpre = gy
            if need_b:
                call('dvsr_act_bwd', _ptr(gy), None, None, None, _ptr(gb), BT * H * W, Co, ACT_NONE, 0.0, 0, 0, H, W, _stream())
        gw = None
        if ctx.needs_input_grad[1]:
            gw = ctx.wslot if ctx.wslot is not None else torch.zeros_like(weight)
            d = _new_desc(weight)
            d.N, d.H, d.W, d.Ho, d.Wo = BT, H + 2, W + 2, H, W
            d.KH, d.KW, d.stride, d.pad, d.dil = 3, 3, 1, 0, 1
            d.nseg = 1
            _fill_seg(d.seg[0], xc)
            d.Co = Co
            if _lib.lib().dvsr_conv_wgrad_tc_supported(ctypes.byref(d), 0) != 1:
                raise RuntimeError('conv3d_rgb: weight gradient shape not supported')
            _run_wgrad(d, gpre, Co, gw, ctx.wl, keep=(gpre, gy, xc) if ctx.wslot is not None else None, force_tc=True, weight=weight)
        return None, (None if ctx.wslot is not None else gw), (None if ctx.bslot is not None else gb), None, None, None


def conv3d_rgb(x, weight, bias, T, act=ACT_NONE, slope=0.1):
    """Conv3d(C <= 4 -> Co, 3^3, replication padding 1) on clip frames [B*T, H, W, C] -> [B*T, H, W, Co] (tensor cores)."""
    return _Conv3dRgbFn.apply(x, weight, bias, T, act, slope)


# --------------------------------------------------------------------------------------------------
# modulated deformable convolution (NHWC, fused gather + GEMM)
class _MdcnFn(Function):
    @staticmethod
    def forward(ctx, x, om, weight, bias, dg, stride, pad, dil, act, slope):
        _check_cuda(x, om, weight, bias)
        N, H, W, C = x.shape
        Co, _, KH, KW = weight.shape
        KK = KH * KW
        Ho = (H + 2 * pad - (dil * (KH - 1) + 1)) // stride + 1
        Wo = (W + 2 * pad - (dil * (KW - 1) + 1)) // stride + 1
        assert om.shape == (N, Ho, Wo, 3 * dg * KK) and om.is_contiguous() and x.is_contiguous()
        wl = _layout(weight, [C], False)
        d = _MdcnFn._desc(x, om, dg, KH, KW, stride, pad, dil, Ho, Wo, weight)
        d.Co = Co
        d.bias = bias.data_ptr() if bias is not None else None
        d.act, d.slope = act, slope
        y = torch.empty(N, Ho, Wo, Co, device=x.device, dtype=torch.float32)
        d.y, d.y_pix_stride = y.data_ptr(), Co
        if _lib.PROFILE['on']:
            _lib.PROFILE['tag'] = 'mdcn fwd %dx%dx%d C%d->%d' % (N, Ho, Wo, C, Co)
        if _tc() and _lib.lib().dvsr_mdcn_tc_supported(ctypes.byref(d)) == 1:
            nblocks = KK * ((C + 31) // 32)
            wp = _get_pack(weight, wl, 7, 0, 1, a=(nblocks, 0, 0, 0),
                           total=_lib.lib().dvsr_conv_tc2_packed_floats(ctypes.byref(wl), 7, 0, 1))
            call('dvsr_mdcn_tc_fprop', ctypes.byref(d), _ptr(wp), _stream())
        else:
            call('dvsr_conv_fprop', ctypes.byref(d), _ptr(_packed(weight, wl, 0)), _stream())
        ctx.cfg = (dg, stride, pad, dil, act, slope, Ho, Wo, bias is not None)
        ctx.wl = wl
        ctx.wslot, ctx.bslot = getattr(weight, '_dvsr_grad', None), getattr(bias, '_dvsr_grad', None)
        ctx.save_for_backward(x, om, weight, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def _desc(x, om, dg, KH, KW, stride, pad, dil, Ho, Wo, weight=None):
        N, H, W, C = x.shape
        KK = KH * KW
        d = _new_desc(weight)
        d.N, d.H, d.W, d.Ho, d.Wo = N, H, W, Ho, Wo
        d.KH, d.KW, d.stride, d.pad, d.dil = KH, KW, stride, pad, dil
        d.nseg = 1
        _fill_seg(d.seg[0], x)
        d.deform, d.dg = 1, dg
        d.offset, d.off_pix_stride = om.data_ptr(), om.shape[3]
        d.mask, d.mask_pix_stride = om.data_ptr() + 4 * 2 * dg * KK, om.shape[3]
        return d

    @staticmethod
    def backward(ctx, gy):
        x, om, weight, y = ctx.saved_tensors
        dg, stride, pad, dil, act, slope, Ho, Wo, has_bias = ctx.cfg
        wl = ctx.wl
        N, H, W, C = x.shape
        Co, _, KH, KW = weight.shape
        KK = KH * KW
        gy = gy.contiguous()
        npix = N * Ho * Wo
        need_b = has_bias and ctx.needs_input_grad[3]
        gb = (ctx.bslot if ctx.bslot is not None else torch.zeros(Co, device=gy.device, dtype=torch.float32)) \
            if need_b else None
        if act != ACT_NONE:
            gpre = torch.empty_like(gy)
            call('dvsr_act_bwd', _ptr(gy), _ptr(y), None, _ptr(gpre), _ptr(gb), npix, Co, act, slope, 0, 0, Ho, Wo, _stream())
        else:
            gpre = gy
            if need_b:
                call('dvsr_act_bwd', _ptr(gy), None, None, None, _ptr(gb), npix, Co, ACT_NONE, 0.0, 0, 0, Ho, Wo, _stream())
        d = _MdcnFn._desc(x, om, dg, KH, KW, stride, pad, dil, Ho, Wo, weight)
        d.Co = Co
        gx = gom = gw = None
        need_data = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        if _tc() and _backend.get('mdcn_bwd_tc', True) and _lib.lib().dvsr_mdcn_bwd_tc_supported(ctypes.byref(d)) == 1:
            # ONE tcgen05 kernel: grad_col = gy . W^T in TMEM, offset / mask / input gradients, and the weight gradient from the
            # modulated samples it rebuilds on the way (no `columns` buffer, no second gather pass)
            gx = torch.zeros_like(x) if ctx.needs_input_grad[0] else None
            gom = torch.empty_like(om) if need_data else None
            if ctx.needs_input_grad[2]:
                gw = ctx.wslot if ctx.wslot is not None else torch.zeros_like(weight)
            wp = _get_pack(weight, wl, 10, 0, 1, a=(KK * ((Co + 63) // 64), 0, 0, 0),
                           total=_lib.lib().dvsr_conv_tc2_packed_floats(ctypes.byref(wl), 10, 0, 1))
            if _lib.PROFILE['on']:
                _lib.PROFILE['tag'] = 'mdcn bwd (tc) %dx%dx%d C%d->%d' % (N, Ho, Wo, C, Co)
            call('dvsr_mdcn_bwd_tc', ctypes.byref(d), _ptr(gpre), Co, _ptr(wp), _ptr(gx), C,
                 ctypes.c_void_p(gom.data_ptr()) if gom is not None else None, om.shape[3],
                 ctypes.c_void_p(gom.data_ptr() + 4 * 2 * dg * KK) if gom is not None else None, om.shape[3],
                 _ptr(gw), ctypes.byref(wl), _stream())
        else:
            if need_data:
                gx = torch.zeros_like(x) if ctx.needs_input_grad[0] else None
                gom = torch.empty_like(om)
                goff_p = gom.data_ptr()
                gmask_p = gom.data_ptr() + 4 * 2 * dg * KK
                call('dvsr_mdcn_bwd_data', ctypes.byref(d), _ptr(gpre), Co, _ptr(_packed(weight, wl, 1, 0)),
                     _ptr(gx), C, ctypes.c_void_p(goff_p), om.shape[3], ctypes.c_void_p(gmask_p), om.shape[3], _stream())
            if ctx.needs_input_grad[2]:
                gw = ctx.wslot if ctx.wslot is not None else torch.zeros_like(weight)
                _run_wgrad(d, gpre, Co, gw, wl, keep=(gpre, gy, x, om) if ctx.wslot is not None else None, weight=weight)
        if ctx.wslot is not None:
            gw = None
        if ctx.bslot is not None:
            gb = None
        return gx, gom, gw, gb, None, None, None, None, None, None


def mdcn(x, om, weight, bias=None, deformable_groups=1, stride=1, pad=1, dil=1, act=ACT_NONE, slope=0.1):
    """Modulated deformable conv on NHWC tensors.  ``om`` = [N, Ho, Wo, 3*dg*kh*kw]: offsets
    ([dg][k][dy,dx], the reference layout) in the first 2*dg*k channels, post-sigmoid mask in the rest."""
    return _MdcnFn.apply(x, om, weight, bias, deformable_groups, stride, pad, dil, act, slope)


# --------------------------------------------------------------------------------------------------
# resampling / pooling / padding
class _UpsampleFn(Function):
    @staticmethod
    def forward(ctx, x, scale, mul):
        _check_cuda(x)
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty(N, H * scale, W * scale, C, device=x.device, dtype=torch.float32)
        call('dvsr_upsample_bilinear', _ptr(x), _ptr(y), N, H, W, C, scale, mul, 0, _stream())
        ctx.cfg = (N, H, W, C, scale, mul)
        return y

    @staticmethod
    def backward(ctx, gy):
        N, H, W, C, scale, mul = ctx.cfg
        gy = gy.contiguous()
        gx = torch.empty(N, H, W, C, device=gy.device, dtype=torch.float32)
        call('dvsr_upsample_bilinear_bwd', _ptr(gy), _ptr(gx), N, H, W, C, scale, mul, _stream())
        return gx, None, None


def upsample(x, scale=2, mul=1.0):
    """mul * F.interpolate(x, scale_factor=scale, mode='bilinear', align_corners=False) on NHWC."""
    return _UpsampleFn.apply(x, scale, float(mul))


class _PoolFn(Function):
    @staticmethod
    def forward(ctx, x):
        _check_cuda(x)
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty(N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 2 * C, device=x.device, dtype=torch.float32)
        call('dvsr_pool_maxavg', _ptr(x), _ptr(y), N, H, W, C, _stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, = ctx.saved_tensors
        N, H, W, C = x.shape
        gx = torch.empty_like(x)
        call('dvsr_pool_maxavg_bwd', _ptr(x), _ptr(gy.contiguous()), _ptr(gx), N, H, W, C, _stream())
        return gx


def pool_maxavg(x):
    """cat([max_pool2d(x,3,2,1), avg_pool2d(x,3,2,1)], C) in one pass (EDVR_arch.py:184-186,190-192)."""
    return _PoolFn.apply(x)


class _Pad2dFn(Function):
    @staticmethod
    def forward(ctx, x, p, mode):
        _check_cuda(x)
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty(N, H + 2 * p, W + 2 * p, C, device=x.device, dtype=torch.float32)
        call('dvsr_pad2d', _ptr(x), _ptr(y), N, H, W, C, p, mode, _stream())
        ctx.cfg = (N, H, W, C, p, mode)
        return y

    @staticmethod
    def backward(ctx, gy):
        N, H, W, C, p, mode = ctx.cfg
        gx = torch.empty(N, H, W, C, device=gy.device, dtype=torch.float32)
        call('dvsr_pad2d_bwd', _ptr(gy.contiguous()), _ptr(gx), N, H, W, C, p, mode, _stream())
        return gx, None, None


def pad2d(x, p=1, mode='reflect'):
    return _Pad2dFn.apply(x, p, 0 if mode == 'reflect' else 1)


class _Pad3dFn(Function):
    @staticmethod
    def forward(ctx, x, T):
        _check_cuda(x)
        x = x.contiguous()
        BT, H, W, C = x.shape
        B = BT // T
        y = torch.empty(B * (T + 2), H + 2, W + 2, C, device=x.device, dtype=torch.float32)
        call('dvsr_pad3d_replicate', _ptr(x), _ptr(y), B, T, H, W, C, _stream())
        ctx.cfg = (B, T, H, W, C)
        return y

    @staticmethod
    def backward(ctx, gy):
        B, T, H, W, C = ctx.cfg
        gx = torch.empty(B * T, H, W, C, device=gy.device, dtype=torch.float32)
        call('dvsr_pad3d_replicate_bwd', _ptr(gy.contiguous()), _ptr(gx), B, T, H, W, C, _stream())
        return gx, None


def pad3d_replicate(x, T):
    """nn.ReplicationPad3d(1) on a clip stored as frames [B*T, H, W, C] -> [B*(T+2), H+2, W+2, C]."""
    return _Pad3dFn.apply(x, T)


# --------------------------------------------------------------------------------------------------
# TSA
class _TsaTemporalFn(Function):
    @staticmethod
    def forward(ctx, aligned, emb, emb_ref, F_):
        _check_cuda(aligned, emb, emb_ref)
        BF, H, W, C = aligned.shape
        B = BF // F_
        aligned, emb, emb_ref = aligned.contiguous(), emb.contiguous(), emb_ref.contiguous()
        prob = torch.empty(B, F_, H, W, device=aligned.device, dtype=torch.float32)
        out = torch.empty(B, H, W, F_ * C, device=aligned.device, dtype=torch.float32)
        call('dvsr_tsa_temporal', _ptr(aligned), _ptr(emb), _ptr(emb_ref), _ptr(prob), _ptr(out), B, F_, H * W, C, _stream())
        ctx.F = F_
        ctx.save_for_backward(aligned, emb, emb_ref, prob)
        return out

    @staticmethod
    def backward(ctx, gout):
        aligned, emb, emb_ref, prob = ctx.saved_tensors
        BF, H, W, C = aligned.shape
        B = BF // ctx.F
        ga, ge, gr = torch.empty_like(aligned), torch.empty_like(emb), torch.empty_like(emb_ref)
        call('dvsr_tsa_temporal_bwd', _ptr(aligned), _ptr(emb), _ptr(emb_ref), _ptr(prob), _ptr(gout.contiguous()),
             _ptr(ga), _ptr(ge), _ptr(gr), B, ctx.F, H * W, C, _stream())
        return ga, ge, gr, None


def tsa_temporal(aligned, emb, emb_ref, nframes):
    """aligned_f * sigmoid(sum_c emb_f * emb_ref) for every frame f (EDVR_arch.py:166-176).
    aligned, emb: [B*F, H, W, C]; emb_ref: [B, H, W, C]; returns [B, H, W, F*C] (= aligned_fea.view(B,-1,H,W))."""
    return _TsaTemporalFn.apply(aligned, emb, emb_ref, nframes)


class _TsaCombineFn(Function):
    @staticmethod
    def forward(ctx, fea, att, att_add):
        _check_cuda(fea, att, att_add)
        fea, att, att_add = fea.contiguous(), att.contiguous(), att_add.contiguous()
        out = torch.empty_like(fea)
        call('dvsr_tsa_combine', _ptr(fea), _ptr(att), _ptr(att_add), _ptr(out), fea.numel(), _stream())
        ctx.save_for_backward(fea, att)
        return out

    @staticmethod
    def backward(ctx, gout):
        fea, att = ctx.saved_tensors
        gout = gout.contiguous()
        gf, ga = torch.empty_like(fea), torch.empty_like(att)
        call('dvsr_tsa_combine_bwd', _ptr(fea), _ptr(att), _ptr(gout), _ptr(gf), _ptr(ga), fea.numel(), _stream())
        return gf, ga, gout


def tsa_combine(fea, att, att_add):
    """fea * sigmoid(att) * 2 + att_add (EDVR_arch.py:200-202)."""
    return _TsaCombineFn.apply(fea, att, att_add)


# --------------------------------------------------------------------------------------------------
# losses
_LOSS_KIND = {'l1': _lib.LOSS_L1, 'l2': _lib.LOSS_L2, 'cb': _lib.LOSS_CB, 'huber': _lib.LOSS_HUBER}


class _LossFn(Function):
    @staticmethod
    def forward(ctx, a, b, kind, weight, eps):
        _check_cuda(a, b)
        a, b = a.contiguous(), b.contiguous()
        assert a.shape == b.shape
        loss = torch.zeros((), device=a.device, dtype=torch.float32)
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        call('dvsr_loss_fwd', _ptr(a), _ptr(b), _ptr(loss), _ptr(ga), a.numel(), kind, weight, eps, _stream())
        ctx.save_for_backward(ga)
        return loss

    @staticmethod
    def backward(ctx, g):
        ga, = ctx.saved_tensors
        if ga is None:
            return None, None, None, None, None
        out = torch.empty_like(ga)
        call('dvsr_scale_by_device_scalar', _ptr(ga), _ptr(g.contiguous()), _ptr(out), ga.numel(), _stream())
        return out, None, None, None, None


def pixel_loss(a, b, kind='l2', weight=1.0, eps=1e-6):
    """weight * mean(f(a - b)), f in {l1, l2, Charbonnier (eps), Huber (delta = eps)}; gradient flows to ``a`` only
    (Video_base_model.py:39-50, loss.py:19-30, test_dynavsr.py:274)."""
    return _LossFn.apply(a, b, _LOSS_KIND[kind], float(weight), float(eps))


# --------------------------------------------------------------------------------------------------
# layout at the NCHW boundary
class _ToNhwcFn(Function):
    @staticmethod
    def forward(ctx, x):
        _check_cuda(x)
        x = x.contiguous()
        N, C, H, W = x.shape
        y = torch.empty(N, H, W, C, device=x.device, dtype=torch.float32)
        call('dvsr_nchw_to_nhwc', _ptr(x), _ptr(y), N, C, H, W, _stream())
        return y

    @staticmethod
    def backward(ctx, gy):
        return _ToNchwFn.apply(gy)


class _ToNchwFn(Function):
    @staticmethod
    def forward(ctx, x):
        _check_cuda(x)
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty(N, C, H, W, device=x.device, dtype=torch.float32)
        call('dvsr_nhwc_to_nchw', _ptr(x), _ptr(y), N, C, H, W, _stream())
        return y

    @staticmethod
    def backward(ctx, gy):
        return _ToNhwcFn.apply(gy)


def to_nhwc(x):
    return _ToNhwcFn.apply(x)


def to_nchw(x):
    return _ToNchwFn.apply(x)


def abs_sum(x, c0, c1, out):
    """out[0] += sum |x[..., c0:c1]| (device-side accumulation of the offset-magnitude check)."""
    N, H, W, C = x.shape
    call('dvsr_abs_sum', _ptr(x), _ptr(out), N * H * W, C, c0, c1, _stream())


def frame_to_u8(frame, out=None, ref=None, sse=None, bgr=False):
    """Result frame -> 8-bit HWC image on the device (utils/util.py:112-142 tensor2img: clamp, * 255, round half to even).
    ``frame``: [..., H, W, C] channels-last float32; ``out`` / ``ref``: uint8 tensors of the same shape; with ``ref`` and
    ``sse`` (a one-element int64 device tensor the caller zeroed) the exact sum of squared differences between the new image
    and ``ref`` is ADDED to ``sse`` -- the integer part of calculate_psnr (:262-269).  ``bgr`` reverses the channel order
    (tensor2img's default mode; cv2.imwrite wants it).  Not differentiable."""
    _check_cuda(frame)
    frame = frame.detach().contiguous()
    C = frame.shape[-1]
    if out is None:
        out = torch.empty(frame.shape, dtype=torch.uint8, device=frame.device)
    for t in (out, ref):
        if t is not None and (t.dtype != torch.uint8 or not t.is_cuda or not t.is_contiguous() or t.numel() != frame.numel()):
            raise RuntimeError('frame_to_u8: image buffers must be contiguous CUDA uint8 tensors with the shape of the frame')
    if sse is not None and (sse.dtype != torch.int64 or not sse.is_cuda or sse.numel() != 1):
        raise RuntimeError('frame_to_u8: sse must be a one-element int64 CUDA tensor')
    call('dvsr_frame_to_u8', _ptr(frame), _ptr(out), _ptr(ref), _ptr(sse), frame.numel() // C, C, 1 if bgr else 0, _stream())
    return out
