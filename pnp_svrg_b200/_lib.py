"""ctypes binding of libpnp_b200.so (C ABI in include/pnp_b200.h).

There is NO fallback: if the shared library is missing or cannot be loaded the import of any
compute path raises.  Build it with ``python -c 'import __graft_entry__ as g; g.build()'`` or
``python -m pnp_svrg_b200.build``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PNP_LIB selects another build of the same library (e.g. the -DPNP_TRACE build used by scripts/trace_iter.py)
LIB_PATH = os.environ.get('PNP_LIB') or os.path.join(_HERE, 'lib', 'libpnp_b200.so')

c_float_p = C.c_void_p      # all device pointers travel as integers
c_int_p = C.c_void_p


class CsmriGradArgs(C.Structure):
    """mirror of pnp_csmri_grad_args"""
    _fields_ = [
        ('H', C.c_int), ('W', C.c_int), ('batch', C.c_int),
        ('a', C.c_void_p), ('b', C.c_void_p), ('S', C.c_void_p), ('bits', C.c_void_p),
        ('Y1', C.c_void_p), ('Y2', C.c_void_p), ('Y1n', C.c_void_p), ('Y2n', C.c_void_p),
        ('gscale', C.c_float), ('gscale_ptr', C.c_void_p),
        ('step', C.c_float), ('step_ptr', C.c_void_p),
        ('g_out', C.c_void_p), ('vadd', C.c_void_p), ('v_out', C.c_void_p),
        ('z_in', C.c_void_p), ('z_out', C.c_void_p),
        ('phases', C.c_int), ('clear_bits', C.c_int),
        ('sel_count', C.c_int), ('sel_idx', C.c_void_p), ('sel_idx_img_stride', C.c_longlong), ('sel_cursor', C.c_void_p),
        ('sel_support', C.c_void_p), ('sel_m0', C.c_void_p), ('sel_support_img_stride', C.c_longlong),
        ('sel_seed', C.c_uint), ('sel_counter', C.c_void_p), ('sel_min_m0', C.c_int), ('flags', C.c_int),
        ('row_lo', C.c_int), ('row_hi', C.c_int),
    ]


class CsmriBuildArgs(C.Structure):
    """mirror of pnp_csmri_build_args"""
    _fields_ = [
        ('H', C.c_int), ('W', C.c_int), ('batch', C.c_int), ('seed', C.c_uint),
        ('x', C.c_void_p), ('p', C.c_void_p), ('snr', C.c_void_p), ('bits_full', C.c_void_p), ('m0', C.c_void_p),
        ('inv_m0', C.c_void_p), ('support', C.c_void_p), ('support_img_stride', C.c_longlong),
        ('Y1', C.c_void_p), ('Y2', C.c_void_p), ('Y1n', C.c_void_p), ('Y2n', C.c_void_p),
        ('xinit', C.c_void_p), ('sigma', C.c_void_p), ('work', C.c_void_p),
    ]


class DeblurGradArgs(C.Structure):
    """mirror of pnp_deblur_grad_args"""
    _fields_ = [
        ('H', C.c_int), ('W', C.c_int), ('batch', C.c_int),
        ('a', C.c_void_p), ('b', C.c_void_p), ('S', C.c_void_p), ('blurred', C.c_void_p), ('up', C.c_void_p),
        ('Bf', C.c_void_p), ('twn', C.c_void_p), ('y', C.c_void_p), ('tl', C.c_void_p), ('wts', C.c_void_p),
        ('identity', C.c_int), ('M', C.c_int), ('sel', C.c_void_p), ('count', C.c_int), ('cursor', C.c_void_p),
        ('use_y', C.c_int), ('gscale', C.c_float), ('step', C.c_float), ('step_ptr', C.c_void_p),
        ('g_out', C.c_void_p), ('vadd', C.c_void_p), ('v_out', C.c_void_p), ('z_in', C.c_void_p), ('z_out', C.c_void_p),
        ('ntaps', C.c_int), ('tap_pos', C.c_void_p), ('tap_w', C.c_void_p),
    ]


class PrGradArgs(C.Structure):
    """mirror of pnp_pr_grad_args"""
    _fields_ = [
        ('A', C.c_void_p), ('n', C.c_longlong), ('M', C.c_int), ('z', C.c_void_p), ('w', C.c_void_p), ('y', C.c_void_p),
        ('rows', C.c_void_p), ('count', C.c_int), ('cursor', C.c_void_p), ('r', C.c_void_p),
        ('gscale', C.c_float), ('step', C.c_float), ('step_ptr', C.c_void_p),
        ('g_out', C.c_void_p), ('vadd', C.c_void_p), ('v_out', C.c_void_p), ('z_in', C.c_void_p), ('z_out', C.c_void_p),
        ('partial', C.c_void_p), ('partial_chunks', C.c_int),
    ]


class CdpGradArgs(C.Structure):
    """mirror of pnp_cdp_grad_args"""
    _fields_ = [
        ('H', C.c_int), ('W', C.c_int), ('L', C.c_int), ('codes', C.c_void_p), ('y', C.c_void_p), ('z', C.c_void_p),
        ('w', C.c_void_p), ('sel_idx', C.c_void_p), ('count', C.c_int), ('cursor', C.c_void_p), ('mask', C.c_void_p),
        ('S', C.c_void_p), ('acc', C.c_void_p), ('gscale', C.c_float), ('step', C.c_float), ('step_ptr', C.c_void_p),
        ('g_out', C.c_void_p), ('vadd', C.c_void_p), ('v_out', C.c_void_p), ('z_in', C.c_void_p), ('z_out', C.c_void_p),
        ('S2', C.c_void_p),
    ]


class CsmriNextPass(C.Structure):
    """mirror of pnp_csmri_next_pass"""
    _fields_ = [
        ('w', C.c_void_p), ('S_out', C.c_void_p), ('bits', C.c_void_p), ('sel_count', C.c_int), ('sel_idx', C.c_void_p),
        ('sel_support', C.c_void_p), ('sel_m0', C.c_void_p), ('sel_seed', C.c_uint), ('sel_counter', C.c_void_p),
        ('sel_counter_add', C.c_int), ('sel_min_m0', C.c_int),
    ]


class SvrgSmallArgs(C.Structure):
    """mirror of pnp_csmri_svrg_small_args"""
    _fields_ = [
        ('H', C.c_int), ('W', C.c_int), ('batch', C.c_int),
        ('z', C.c_void_p), ('xrec', C.c_void_p),
        ('Y1', C.c_void_p), ('Y2', C.c_void_p), ('Y1n', C.c_void_p), ('Y2n', C.c_void_p),
        ('bits_full', C.c_void_p), ('support', C.c_void_p), ('m0', C.c_void_p), ('support_img_stride', C.c_longlong),
        ('idx', C.c_void_p), ('idx_img_stride', C.c_longlong), ('idx_iter_stride', C.c_longlong),
        ('snap_scale_ptr', C.c_void_p), ('snap_scale', C.c_float),
        ('step', C.c_void_p), ('step_img_stride', C.c_longlong),
        ('sig_log', C.c_void_p), ('mse_log', C.c_void_p), ('slot', C.c_void_p), ('draw_counter', C.c_void_p),
        ('n_inner', C.c_int), ('T2', C.c_int), ('mini_batch_size', C.c_int), ('seed', C.c_uint),
        ('lr_decay', C.c_float), ('sigma_modifier', C.c_float), ('fallback_sigma', C.c_float), ('fallback_decay', C.c_float),
    ]


CNN_MAX_LAYERS = 32


class CnnNet(C.Structure):
    """mirror of pnp_cnn_net"""
    _fields_ = [
        ('n_layers', C.c_int),
        ('w', C.c_void_p * CNN_MAX_LAYERS), ('scale', C.c_void_p * CNN_MAX_LAYERS), ('shift', C.c_void_p * CNN_MAX_LAYERS),
        ('slope', C.c_float * CNN_MAX_LAYERS),
        ('last_bias', C.c_float), ('mode', C.c_int), ('range', C.c_float), ('shift_in', C.c_float),
        ('w_tc', C.c_void_p * CNN_MAX_LAYERS), ('w_tc_lo', C.c_void_p * CNN_MAX_LAYERS),
    ]


# name -> (restype, argtypes); every symbol declared in include/pnp_b200.h
PROTOTYPES = {
    'pnp_init': (C.c_int, []),
    'pnp_last_error': (C.c_char_p, []),
    'pnp_version': (C.c_int, []),
    'pnp_csmri_grad': (C.c_int, [C.POINTER(CsmriGradArgs), C.c_void_p]),
    'pnp_csmri_build_batch_workspace': (C.c_longlong, [C.c_int, C.c_int, C.c_int]),
    'pnp_csmri_build_batch': (C.c_int, [C.POINTER(CsmriBuildArgs), C.c_void_p]),
    'pnp_csmri_sel_from_indices': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                             C.c_longlong, C.c_void_p, C.c_int, C.c_void_p]),
    'pnp_csmri_sel_sample': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_longlong, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p]),
    'pnp_sample_indices': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p]),
    'pnp_sample_indices_host': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_void_p]),
    'pnp_host_draws_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int]),
    'pnp_host_draws_set_device_support': (C.c_int, [C.c_void_p, C.c_void_p]),
    'pnp_host_draws_next': (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    'pnp_host_draws_stage': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    'pnp_host_draws_stage_many': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    'pnp_host_draws_destroy': (C.c_int, [C.c_void_p]),
    'pnp_deblur_grad': (C.c_int, [C.POINTER(DeblurGradArgs), C.c_void_p]),
    'pnp_pr_grad': (C.c_int, [C.POINTER(PrGradArgs), C.c_void_p]),
    'pnp_cdp_grad': (C.c_int, [C.POINTER(CdpGradArgs), C.c_void_p]),
    'pnp_nlm_denoise': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pnp_cnn_forward': (C.c_int, [C.POINTER(CnnNet), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'pnp_csmri_update_prox': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'pnp_csmri_update_prox_next': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(CsmriNextPass), C.c_void_p]),
    'pnp_csmri_update_prox_supported': (C.c_int, [C.c_int, C.c_int]),
    'pnp_tv_chambolle': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_float,
                                   C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pnp_estimate_sigma': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pnp_wavelet_denoise': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float,
                                      C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pnp_prox_wavelet_fused': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pnp_sq_err': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pnp_axpy': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_void_p,
                           C.c_void_p]),
    'pnp_saga_init': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    'pnp_saga_update': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                  C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_float, C.c_void_p,
                                  C.c_void_p]),
    'pnp_csmri_svrg_small': (C.c_int, [C.POINTER(SvrgSmallArgs), C.c_void_p]),
    'pnp_csmri_svrg_small_supported': (C.c_int, [C.c_int, C.c_int]),
    'pnp_csmri_svrg_small_capacity': (C.c_int, [C.c_int, C.c_int]),
    'pnp_advance': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    'pnp_advance_by': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'pnp_advance_scale': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p]),
    'pnp_copy_f32': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    'pnp_debug_set': (C.c_int, [C.c_int, C.c_int]),
    'pnp_debug_read': (C.c_int, [C.c_int, C.c_void_p, C.c_longlong]),
    'pnp_graph_begin': (C.c_int, [C.c_void_p]),
    'pnp_graph_end': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    'pnp_graph_launch': (C.c_int, [C.c_void_p, C.c_void_p]),
    'pnp_graph_destroy': (C.c_int, [C.c_void_p]),
}

_lib = None
_inited_devices = set()


class PnpError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnpError('%s not found: the CUDA extension is not built and there is no CPU fallback '
                       '(run `python -m pnp_svrg_b200.build`)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)        # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PnpError('libpnp_b200: status %d: %s' % (rc, load().pnp_last_error().decode()))


def init_device():
    """pnp_init() on torch's current CUDA device (twiddle table, shared-memory limits)."""
    import torch
    if not torch.cuda.is_available():
        raise PnpError('no CUDA device: pnp_svrg_b200 has no CPU path')
    dev = torch.cuda.current_device()
    if dev not in _inited_devices:
        torch.cuda.init()
        torch.zeros(1, device='cuda')          # make sure the primary context exists and is current
        check(load().pnp_init())
        _inited_devices.add(dev)
    return dev
