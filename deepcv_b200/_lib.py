""" ctypes binding of the C ABI declared in `include/deepcv_b200.h` (built in-tree as `deepcv_b200/libdeepcv_b200.so`).

There is no CPU implementation behind these symbols and no fallback here either: if the shared library is missing, or a
call returns non-zero, a `RuntimeError` is raised with the library's own message.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

__all__ = ['lib', 'ConvShape', 'NormParams', 'ScNorm', 'PackEntry', 'LinkSource', 'LINK_MAX_SOURCES', 'check', 'library_path', 'DCV_F32', 'DCV_BF16', 'ACT_NONE', 'ACT_RELU', 'ACT_LEAKY_RELU',
           'ACT_SIGMOID', 'ALGO_AUTO', 'ALGO_DIRECT', 'ALGO_TCGEN05', 'SYMBOLS']

ABI_VERSION = 4
DCV_F32, DCV_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU, ACT_SIGMOID = 0, 1, 2, 3
ALGO_AUTO, ALGO_DIRECT, ALGO_TCGEN05 = 0, 1, 2


class ConvShape(Structure):
    _fields_ = [(n, c_int32) for n in ('n', 'h', 'w', 'c', 'k', 'r', 's', 'stride_h', 'stride_w', 'pad_h', 'pad_w', 'dil_h', 'dil_w', 'p', 'q')]


class NormParams(Structure):
    _fields_ = [('n', c_int32), ('c', c_int32), ('hw', c_int32), ('use_bn', c_int32), ('bn_training', c_int32),
                ('bn_eps', c_float), ('bn_momentum', c_float),
                ('bn_weight', c_void_p), ('bn_bias', c_void_p), ('bn_running_mean', c_void_p), ('bn_running_var', c_void_p),
                ('bn_num_batches_tracked', c_void_p),
                ('use_gn', c_int32), ('gn_groups', c_int32), ('gn_eps', c_float),
                ('gn_weight', c_void_p), ('gn_bias', c_void_p)]


class PackEntry(Structure):
    """ `dcv_pack_entry` (dcv_pack_conv_weights_batched) """
    _fields_ = [('src_off', c_uint64), ('dst_off', c_uint64), ('unit0', c_int64), ('k', c_int32), ('r', c_int32), ('s', c_int32), ('c', c_int32)]


class LinkSource(Structure):
    """ `dcv_link_source` (dcv_link_concat_fwd / _bwd) """
    _fields_ = [('ptr', c_void_p), ('channels', c_int32), ('pool', c_int32)]


LINK_MAX_SOURCES = 8


class ScNorm(Structure):
    """ `dcv_sc_norm`: a raw block output's pending BatchNorm o GroupNorm, its raw sums and backward sum buffers (few-channel path). """
    _fields_ = [('enabled', c_int32), ('n', c_int32), ('c', c_int32), ('hw', c_int32), ('use_bn', c_int32), ('bn_training', c_int32),
                ('bn_eps', c_float), ('bn_momentum', c_float),
                ('bn_weight', c_void_p), ('bn_bias', c_void_p), ('bn_running_mean', c_void_p), ('bn_running_var', c_void_p), ('bn_num_batches_tracked', c_void_p),
                ('use_gn', c_int32), ('gn_groups', c_int32), ('gn_eps', c_float),
                ('gn_weight', c_void_p), ('gn_bias', c_void_p),
                ('stats_nc', c_void_p), ('bn_sums', c_void_p), ('s_nc', c_void_p), ('u_sums', c_void_p), ('coef_nc', c_void_p), ('d_nc', c_void_p)]


P = c_void_p
""" name -> (restype, argtypes); must list every symbol `include/deepcv_b200.h` declares (tests/test_abi.py checks both ways). """
SYMBOLS = {
    'dcv_abi_version': (c_int, []),
    'dcv_last_error': (c_char_p, []),
    'dcv_device_check': (c_int, []),
    'dcv_launch_count': (c_uint64, []),
    'dcv_preprocess_u8': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, c_int, c_int, c_int, P]),
    'dcv_nchw_to_nhwc': (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_nhwc_to_nchw': (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_cast': (c_int, [P, c_int, P, c_int, c_size_t, P]),
    'dcv_pack_conv_weight': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_im2col': (c_int, [POINTER(ConvShape), P, P, c_int, c_int, P]),
    'dcv_fill_zero': (c_int, [P, c_size_t, P]),
    'dcv_gather_rows': (c_int, [P, P, P, c_int, c_int64, c_size_t, P, P]),
    'dcv_conv2d_tc_supported': (c_int, [POINTER(ConvShape), c_int, c_int]),
    'dcv_conv2d_fwd': (c_int, [POINTER(ConvShape), P, P, P, P, P, c_int, c_float, c_int, c_int, c_int, P]),
    'dcv_conv2d_dgrad': (c_int, [POINTER(ConvShape), P, P, P, P, c_int, c_int, P]),
    'dcv_conv2d_wgrad_workspace': (c_size_t, [POINTER(ConvShape), c_int, c_int]),
    'dcv_conv2d_wgrad': (c_int, [POINTER(ConvShape), P, P, P, P, c_int, c_int, c_int, P]),
    'dcv_sc_conv_supported': (c_int, [POINTER(ConvShape), c_int]),
    'dcv_sc_norm_floats': (c_size_t, [c_int, c_int, c_int]),
    'dcv_sc_conv_fwd': (c_int, [POINTER(ConvShape), P, POINTER(ScNorm), c_int, P, P, c_int, c_float, P, POINTER(ScNorm), P]),
    'dcv_sc_conv_wgrad': (c_int, [POINTER(ConvShape), P, POINTER(ScNorm), P, P, POINTER(ScNorm), c_int, c_float, P, P, P, P, P, P, P]),
    'dcv_sc_conv_dgrad': (c_int, [POINTER(ConvShape), P, P, POINTER(ScNorm), c_int, c_float, P, P, P, POINTER(ScNorm), P]),
    'dcv_sc_conv_bwd_supported': (c_int, [POINTER(ConvShape), c_int]),
    'dcv_sc_conv_bwd': (c_int, [POINTER(ConvShape), P, POINTER(ScNorm), P, P, POINTER(ScNorm), c_int, c_float, P, P, P, P, P, P, P, P, P]),
    'dcv_sc_affine_pool_fwd': (c_int, [P, POINTER(ScNorm), c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_sc_affine_pool_bwd': (c_int, [P, P, POINTER(ScNorm), P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_peer_flag_words': (c_size_t, []),
    'dcv_peer_max_floats': (c_size_t, []),
    'dcv_peer_allreduce_sum': (c_int, [P, P, P, c_int, c_int, c_size_t, c_size_t, c_size_t, c_int, P, P]),
    'dcv_pack_conv_weights_batched': (c_int, [P, P, c_int, P, c_int, c_int64, P]),
    'dcv_dropout': (c_int, [P, P, P, c_size_t, c_float, c_uint64, P, P, c_int, P]),
    'dcv_activation_fwd': (c_int, [P, P, c_size_t, c_int, c_float, c_int, P]),
    'dcv_activation_bwd': (c_int, [P, P, P, c_size_t, c_int, c_float, c_int, P]),
    'dcv_norm_stats': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_norm_saved_floats': (c_size_t, [c_int, c_int, c_int]),
    'dcv_norm_fwd_finalize': (c_int, [POINTER(NormParams), P, P, P, P]),
    'dcv_norm_apply_fwd': (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    'dcv_norm_apply_add_fwd': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dcv_bn_apply_fold_fwd': (c_int, [P, P, c_int, P, P, P, P, P, P, P, c_float, c_float, c_int, P, c_int, c_int, c_int, c_int, P]),
    'dcv_norm_apply_pool_fwd': (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_conv2d_gather_supported': (c_int, [POINTER(ConvShape), P, c_int, c_int]),
    'dcv_gather_pack_weight': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_gather_unpack_wgrad': (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    'dcv_conv2d_fwd_gather': (c_int, [POINTER(ConvShape), P, P, c_int, P, P, P, c_int, c_float, c_int, P]),
    'dcv_conv2d_wgrad_gather': (c_int, [POINTER(ConvShape), P, P, P, c_int, c_int, P]),
    'dcv_conv2d_pairs_supported': (c_int, [POINTER(ConvShape), P, c_int]),
    'dcv_pairs_pack_weight': (c_int, [P, P, POINTER(ConvShape), c_int, P]),
    'dcv_pairs_unpack_wgrad': (c_int, [P, P, POINTER(ConvShape), P]),
    'dcv_conv2d_fwd_pairs': (c_int, [POINTER(ConvShape), P, P, P, P, P, c_int, c_float, c_int, P]),
    'dcv_conv2d_wgrad_pairs': (c_int, [POINTER(ConvShape), P, P, P, c_int, P]),
    'dcv_norm_bwd_reduce': (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_norm_bwd_finalize': (c_int, [POINTER(NormParams), P, P, P, P, P, P, P, P, P]),
    'dcv_act_norm_bwd_apply': (c_int, [P, P, P, P, P, c_int, c_float, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_norm_bwd_pooled_supported': (c_int, [c_int, c_int, c_int, c_int, c_int]),
    'dcv_act_bn_bwd_apply_fold': (c_int, [P, c_int, P, P, P, c_int, P, P, P, P, c_int, c_float, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_norm_bwd_reduce_pooled': (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_act_norm_bwd_apply_pooled': (c_int, [P, P, P, P, P, c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_avgpool2d_fwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_avgpool2d_bwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_axpby': (c_int, [P, P, P, c_float, c_float, c_size_t, c_int, P]),
    'dcv_link_concat_fwd': (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, P]),
    'dcv_link_concat_bwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_copy_channels_in': (c_int, [P, P, c_size_t, c_int, c_int, c_int, c_int, P]),
    'dcv_copy_channels_out': (c_int, [P, P, c_size_t, c_int, c_int, c_int, c_int, P]),
    'dcv_bilinear_fwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_bilinear_bwd': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    'dcv_linear_flatten_fused': (c_int, [c_int, c_int]),
    'dcv_linear_fwd': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, P]),
    'dcv_linear_bwd': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_int, P]),
    'dcv_softmax_ce': (c_int, [P, P, P, P, c_int, c_int, P]),
    'dcv_classification_metrics': (c_int, [P, P, P, c_int, c_int, P]),
    'dcv_scale_by_device_scalar': (c_int, [P, P, P, c_size_t, P]),
    'dcv_counter_add': (c_int, [P, c_int32, P]),
    'dcv_adamw_flat': (c_int, [P, P, P, P, c_size_t, P, c_float, c_float, c_float, c_float, c_float, P, P]),
}


def library_path() -> Path:
    return Path(os.environ.get('DEEPCV_B200_LIB', Path(__file__).resolve().parent / 'libdeepcv_b200.so'))


class _Library:
    """ Lazy handle: the shared object is loaded on first attribute access so that importing the package (spec parsing,
    shape inference on `meta` tensors) works in a process that never launches a kernel. """

    def __init__(self):
        self._cdll = None

    def _load(self):
        if self._cdll is None:
            path = library_path()
            if not path.exists():
                raise RuntimeError(f'deepcv_b200: CUDA extension "{path}" is missing. Build it with `python -m deepcv_b200.csrc.build` '
                                   '(or `__graft_entry__.build()`); there is no CPU or PyTorch fallback for this path.')
            cdll = ctypes.CDLL(str(path))
            for name, (restype, argtypes) in SYMBOLS.items():
                fn = getattr(cdll, name)
                fn.restype, fn.argtypes = restype, argtypes
            if cdll.dcv_abi_version() != ABI_VERSION:
                raise RuntimeError(f'deepcv_b200: ABI version mismatch ({cdll.dcv_abi_version()} != {ABI_VERSION}); rebuild the extension')
            self._cdll = cdll
        return self._cdll

    def __getattr__(self, name):
        return getattr(self._load(), name)


lib = _Library()


def check(status: int, what: str = '') -> None:
    if status != 0:
        msg = lib.dcv_last_error()
        raise RuntimeError(f'deepcv_b200{": " + what if what else ""}: {msg.decode() if msg else "error " + str(status)}')
