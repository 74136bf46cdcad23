"""ctypes binding of libdvsr_b200.so (the C ABI declared in include/dvsr_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing (or a kernel reports an
error) the call raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C dynavsr_b200/csrc``.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdvsr_b200.so')
MAX_SEG = 5

c_float_p = ctypes.c_void_p  # raw device pointers are passed as integers


class ConvSeg(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('C', ctypes.c_int), ('pix_stride', ctypes.c_int),
                ('img_stride', ctypes.c_longlong), ('T', ctypes.c_int), ('Tsrc', ctypes.c_int),
                ('dt', ctypes.c_int), ('t_fixed', ctypes.c_int)]


class Policy(ctypes.Structure):
    """dvsr_policy: launch policy of one call (all zero = library defaults)."""
    _fields_ = [('cta_budget', ctypes.c_int), ('min_tiles', ctypes.c_int), ('min_chunks', ctypes.c_int),
                ('precision', ctypes.c_int), ('mdcn_staged', ctypes.c_int)]


PREC_BF16X3, PREC_TF32, PREC_BF16 = 0, 1, 2


class ConvDesc(ctypes.Structure):
    _fields_ = [('N', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int),
                ('Ho', ctypes.c_int), ('Wo', ctypes.c_int),
                ('KH', ctypes.c_int), ('KW', ctypes.c_int), ('stride', ctypes.c_int), ('pad', ctypes.c_int),
                ('dil', ctypes.c_int), ('transposed', ctypes.c_int), ('wshare', ctypes.c_int),
                ('nseg', ctypes.c_int), ('seg', ConvSeg * MAX_SEG), ('Co', ctypes.c_int),
                ('deform', ctypes.c_int), ('dg', ctypes.c_int),
                ('offset', ctypes.c_void_p), ('off_pix_stride', ctypes.c_int),
                ('mask', ctypes.c_void_p), ('mask_pix_stride', ctypes.c_int),
                ('bias', ctypes.c_void_p), ('act', ctypes.c_int), ('slope', ctypes.c_float),
                ('sig_split', ctypes.c_int), ('res', ctypes.c_void_p), ('res_pix_stride', ctypes.c_int),
                ('shuffle', ctypes.c_int), ('accumulate', ctypes.c_int),
                ('y', ctypes.c_void_p), ('y_pix_stride', ctypes.c_int),
                ('out_step', ctypes.c_int), ('out_off_y', ctypes.c_int), ('out_off_x', ctypes.c_int),
                ('out_H', ctypes.c_int), ('out_W', ctypes.c_int), ('policy', Policy)]


class WLayout(ctypes.Structure):
    _fields_ = [('co_stride', ctypes.c_longlong), ('ci_stride', ctypes.c_longlong),
                ('seg_base', ctypes.c_longlong * MAX_SEG), ('seg_C', ctypes.c_int * MAX_SEG),
                ('nseg', ctypes.c_int), ('taps', ctypes.c_int), ('Co', ctypes.c_int),
                ('ci_bits', ctypes.c_int), ('ci_lo_valid', ctypes.c_int), ('ci_hi_stride', ctypes.c_longlong)]


class PackJob(ctypes.Structure):
    _fields_ = [('w', ctypes.c_void_p), ('wp', ctypes.c_void_p), ('wl', WLayout),
                ('mode', ctypes.c_int), ('seg', ctypes.c_int), ('seg_hi', ctypes.c_int),
                ('a0', ctypes.c_int), ('a1', ctypes.c_int), ('a2', ctypes.c_int), ('a3', ctypes.c_int),
                ('total', ctypes.c_longlong), ('block_start', ctypes.c_longlong)]


ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID_SPLIT = 0, 1, 2, 3
LOSS_L1, LOSS_L2, LOSS_CB, LOSS_HUBER = 0, 1, 2, 3

_I, _LL, _F, _P = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
_DP, _WP = ctypes.POINTER(ConvDesc), ctypes.POINTER(WLayout)

# name -> argtypes (restype is int unless listed in _RESTYPE).  Kept in sync with include/dvsr_b200.h;
# tests/test_abi.py checks that every symbol declared in the header is exported and listed here.
SIGNATURES = {
    'dvsr_pack_weights': [_P, _P, _WP, _I, _I, _P],
    'dvsr_pack_job_run': [ctypes.POINTER(PackJob), _P],
    'dvsr_pack_table': [_P, _I, _LL, _P],
    'dvsr_pack_table_copy': [_P, _I, _LL, _P, _I, _P],
    'dvsr_conv_fprop': [_DP, _P, _P],
    'dvsr_conv_wgrad': [_DP, _P, _I, _P, _WP, _P],
    'dvsr_conv_small_co': [_DP, _P, _P],
    'dvsr_conv_tc_supported': [_DP],
    'dvsr_conv_tc_packed_floats': [_WP, _I, _I],
    'dvsr_pack_weights_tc': [_P, _P, _WP, _I, _I, _P],
    'dvsr_conv_tc_fprop': [_DP, _P, _P],
    'dvsr_pack_weights_tc_parity': [_P, _P, _WP, _I, _I, _I, _I, _I, _P],
    'dvsr_conv_tc2_supported': [_DP],
    'dvsr_conv_tc2_packed_floats': [_WP, _I, _I, _I],
    'dvsr_pack_weights_tc2': [_P, _P, _WP, _I, _I, _I, _P],
    'dvsr_conv_tc2_fprop': [_DP, _P, _P, _I, _P],
    'dvsr_conv_tc2_set_trace': [_P],
    'dvsr_conv_wgrad_tc_supported': [_DP, _I],
    'dvsr_conv_wgrad_tc': [_DP, _I, _P, _I, _P, _WP, _P],
    'dvsr_mdcn_bwd_data': [_DP, _P, _I, _P, _P, _I, _P, _I, _P, _I, _P],
    'dvsr_mdcn_bwd_tc_supported': [_DP],
    'dvsr_mdcn_bwd_tc': [_DP, _P, _I, _P, _P, _I, _P, _I, _P, _I, _P, _WP, _P],
    'dvsr_mdcn_tc_supported': [_DP],
    'dvsr_mdcn_tc_fprop': [_DP, _P, _P],
    'dvsr_mdcn_workspace_bytes': [_I] * 15,
    'dvsr_mdcn_forward_nchw': [_P] * 6 + [_I] * 15 + [_P, _LL, _P],
    'dvsr_mdcn_backward_nchw': [_P] * 10 + [_I] * 15 + [_P, _LL, _P],
    'dvsr_nchw_to_nhwc': [_P, _P, _I, _I, _I, _I, _P],
    'dvsr_nhwc_to_nchw': [_P, _P, _I, _I, _I, _I, _P],
    'dvsr_upsample_bilinear': [_P, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    'dvsr_upsample_bilinear_bwd': [_P, _P, _I, _I, _I, _I, _I, _F, _P],
    'dvsr_pool_maxavg': [_P, _P, _I, _I, _I, _I, _P],
    'dvsr_pool_maxavg_bwd': [_P, _P, _P, _I, _I, _I, _I, _P],
    'dvsr_pad2d': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'dvsr_pad2d_bwd': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'dvsr_pad3d_replicate': [_P, _P, _I, _I, _I, _I, _I, _P],
    'dvsr_pad3d_replicate_bwd': [_P, _P, _I, _I, _I, _I, _I, _P],
    'dvsr_tcat_pad3': [_P, _P, _I, _I, _I, _I, _I, _P],
    'dvsr_degrade': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'dvsr_spatial_mean': [_P, _P, _I, _I, _I, _P],
    'dvsr_add_channel_bias': [_P, _P, _P, _I, _I, _I, _F, _P],
    'dvsr_frame_to_u8': [_P, _P, _P, _P, _LL, _I, _I, _P],
    'dvsr_act_bwd': [_P, _P, _P, _P, _P, _LL, _I, _I, _F, _I, _I, _I, _I, _P],
    'dvsr_tsa_temporal': [_P, _P, _P, _P, _P, _I, _I, _LL, _I, _P],
    'dvsr_tsa_temporal_bwd': [_P] * 8 + [_I, _I, _LL, _I, _P],
    'dvsr_tsa_combine': [_P, _P, _P, _P, _LL, _P],
    'dvsr_tsa_combine_bwd': [_P, _P, _P, _P, _P, _LL, _P],
    'dvsr_loss_fwd': [_P, _P, _P, _P, _LL, _I, _F, _F, _P],
    'dvsr_scale_by_device_scalar': [_P, _P, _P, _LL, _P],
    'dvsr_update_sgd': [_P, _P, _LL, _LL, _F, _F, _F, _P],
    'dvsr_update_adam': [_P, _P, _P, _P, _LL, _LL, _F, _F, _F, _F, _F, _F, _F, _F, _P],
    'dvsr_update_peers': [_P, _P, _I, _LL, _F, _P, _P, _LL, _LL, _F, _F, _F, _F, _F, _F, _F, _F, _I, _P],
    'dvsr_update_peers_sliced': [_P, _P, _I, _I, _LL, _F, _P, _P, _LL, _LL, _F, _F, _F, _F, _F, _F, _F, _F, _I, _P],
    'dvsr_abs_sum': [_P, _P, _LL, _I, _I, _I, _P],
    'dvsr_sm_count': [],
    'dvsr_last_error': [],
    'dvsr_version': [],
}
_RESTYPE = {'dvsr_last_error': ctypes.c_char_p, 'dvsr_mdcn_workspace_bytes': ctypes.c_longlong,
            'dvsr_conv_tc_packed_floats': ctypes.c_longlong, 'dvsr_conv_tc2_packed_floats': ctypes.c_longlong}

_lib = None


class DvsrError(RuntimeError):
    pass


def build(verbose=False):
    """Compile dynavsr_b200/csrc/*.cu for sm_100a into dynavsr_b200/libdvsr_b200.so (nvcc, no GPU needed)."""
    cmd = ['make', '-C', os.path.join(_HERE, 'csrc'), '-j', str(min(8, os.cpu_count() or 1))]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode:
        print(r.stdout)
    if r.returncode:
        raise DvsrError('building libdvsr_b200.so failed')
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DvsrError('%s is missing: the CUDA extension is not built (there is no CPU fallback); '
                            'run `make -C dynavsr_b200/csrc` or __graft_entry__.build()' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, ctypes.c_int)
        _lib = L
    return _lib


COUNTER = [0]   # kernel-launching C-ABI calls issued from this process (bench.py's gpu_launches)
PROFILE = {'on': bool(os.environ.get('DVSR_PROFILE')), 'events': [], 'tag': ''}


def call(name, *args):
    """Invoke an int-returning entry point; raise DvsrError with the library's message on failure."""
    COUNTER[0] += 1
    if PROFILE['on']:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        PROFILE['events'].append((name, PROFILE['tag'], e0, e1))
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        msg = lib().dvsr_last_error()
        raise DvsrError('%s failed (rc=%d): %s' % (name, rc, msg.decode() if msg else '?'))


def profile_report(top=40, by_tag=True):
    """Aggregate the CUDA-event timings collected while PROFILE['on'] (eager mode only)."""
    import torch
    torch.cuda.synchronize()
    agg = {}
    for name, tag, e0, e1 in PROFILE['events']:
        key = (name, tag) if by_tag else (name, '')
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
    total = sum(v[1] for v in agg.values())
    lines = ['%-28s %-44s %6s %10s %7s' % ('entry', 'tag', 'calls', 'total ms', 'share')]
    for (name, tag), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append('%-28s %-44s %6d %10.3f %6.1f%%' % (name, tag, c, ms, 100 * ms / max(total, 1e-9)))
    lines.append('total %.3f ms over %d calls' % (total, len(PROFILE['events'])))
    PROFILE['events'].clear()
    return '\n'.join(lines)
