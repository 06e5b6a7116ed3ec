"""ctypes binding of libsgr.so (include/sgr.h).  There is no CPU fallback: if the library is missing or a call fails,
a RuntimeError is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SGR_LIB') or os.path.join(_HERE, 'libsgr.so')     # SGR_LIB: A/B experiments only

MAX_STYLED = 24
MAX_RGB = 12
STYLE_DIM = 512
FMT_BF16, FMT_FP16 = 0, 1


def default_format():
    """Forward operand format: bf16 hi/lo split (default, fp32 range) or fp16 hi/lo via SGR_PRECISION=fp16x3
    (22-bit operands, statically scaled, saturating; csrc/sgr_ptx.cuh).  Measured on B200 the two agree to within the
    tensor core's fp32-accumulate rounding for K >= 2304, so the range-safe format is the default."""
    v = os.environ.get('SGR_PRECISION', 'bf16x3').lower()
    if v not in ('fp16x3', 'bf16x3', 'bf16', 'bf16x1'):
        raise RuntimeError('SGR_PRECISION must be fp16x3, bf16x3 or bf16 (single-pass), got %r' % v)
    return FMT_FP16 if v == 'fp16x3' else FMT_BF16


def single_pass():
    """SGR_PRECISION=bf16: one bf16 MMA per product (plain tensor-core precision, BASELINE config 5) instead of the
    3-MMA split that reproduces fp32.  Parity is then reported, not gated at 1e-3 (SURVEY.md §8d cfg 5)."""
    return 1 if os.environ.get('SGR_PRECISION', 'bf16x3').lower() in ('bf16', 'bf16x1') else 0

_fp = C.c_void_p          # device pointers travel as integers


class ConvArgs(C.Structure):
    _fields_ = [('batch', C.c_int), ('cin', C.c_int), ('cout', C.c_int), ('h_in', C.c_int), ('w_in', C.c_int),
                ('ksize', C.c_int), ('up', C.c_int), ('act', C.c_int), ('act_gain', C.c_float), ('operand_format', C.c_int), ('column_tile', C.c_int), ('out_format', C.c_int),
                ('x_c8', _fp), ('w_packed', _fp), ('demod', _fp), ('bias', _fp), ('noise', _fp),
                ('noise_batch_stride', C.c_longlong), ('noise_weight', _fp), ('s2', _fp), ('out_c8', _fp),
                ('out_f32', _fp), ('rgb_coef', _fp), ('rgb_partial', _fp), ('t_scratch', _fp), ('fir', _fp), ('splitk_scratch', _fp), ('splitk_scratch_bytes', C.c_size_t),
                ('single_pass', C.c_int)]


class StyledLayer(C.Structure):
    _fields_ = [('cin', C.c_int), ('cout', C.c_int), ('up', C.c_int), ('latent_row', C.c_int),
                ('column_tile', C.c_int), ('column_tile_t', C.c_int), ('w_packed', _fp), ('w_packed_t', _fp), ('wsq', _fp), ('mod_weight', _fp), ('mod_bias', _fp), ('noise', _fp),
                ('noise_batch_stride', C.c_longlong), ('noise_weight', _fp), ('act_bias', _fp), ('fir', _fp)]


class RgbLayer(C.Structure):
    _fields_ = [('cin', C.c_int), ('latent_row', C.c_int), ('weight', _fp), ('mod_weight', _fp), ('mod_bias', _fp),
                ('bias', _fp), ('fir', _fp), ('fir_flipped', _fp)]


class StyledParamGrads(C.Structure):
    _fields_ = [('weight', _fp), ('g_weight', _fp), ('g_mod_weight', _fp), ('g_mod_bias', _fp), ('g_noise_weight', _fp),
                ('g_act_bias', _fp)]


class RgbParamGrads(C.Structure):
    _fields_ = [('g_weight', _fp), ('g_mod_weight', _fp), ('g_mod_bias', _fp), ('g_bias', _fp)]


class ParamGrads(C.Structure):
    _fields_ = [('styled', StyledParamGrads * MAX_STYLED), ('rgb', RgbParamGrads * MAX_RGB), ('g_const_input', _fp)]


class BackwardExtras(C.Structure):
    _fields_ = [('gfeats', C.POINTER(_fp)), ('ds_styled', C.POINTER(_fp)), ('ds_rgb', C.POINTER(_fp)), ('g_input', _fp),
                ('params', C.POINTER(ParamGrads)), ('wgrad_scratch', _fp), ('wgrad_scratch_bytes', C.c_size_t)]


class ForwardExtras(C.Structure):
    _fields_ = [('frames_u8', _fp), ('u8_h', C.c_int), ('u8_w', C.c_int)]


class WgradArgs(C.Structure):
    _fields_ = [('batch', C.c_int), ('cin', C.c_int), ('cout', C.c_int), ('h_in', C.c_int), ('w_in', C.c_int),
                ('up', C.c_int), ('x_c8', _fp), ('gz_c8', _fp), ('gw', _fp), ('scratch', _fp), ('scratch_bytes', C.c_size_t)]


class Synthesis(C.Structure):
    _fields_ = [('size', C.c_int), ('n_styled', C.c_int), ('n_rgb', C.c_int), ('n_latent', C.c_int), ('format', C.c_int),
                ('single_pass', C.c_int), ('const_input', _fp), ('styled', StyledLayer * MAX_STYLED), ('rgb', RgbLayer * MAX_RGB)]


# name -> (restype, argtypes); mirrors include/sgr.h one to one (tests check every symbol is exported)
SIGNATURES = {
    'sgr_version': (C.c_char_p, []),
    'sgr_last_error': (C.c_char_p, []),
    'sgr_launch_count': (C.c_longlong, []),
    'sgr_reset_launch_count': (None, []),
    'sgr_profile_enable': (None, [C.c_int]),
    'sgr_profile_collect': (C.c_int, [C.POINTER(C.c_float), C.c_int]),
    'sgr_profile_collect_tagged': (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    'sgr_upfirdn2d': (C.c_int, [_fp, _fp, _fp] + [C.c_int] * 9 + [_fp]),
    'sgr_fused_bias_act': (C.c_int, [_fp, _fp, _fp, _fp, C.c_longlong, C.c_int, C.c_longlong, C.c_int, C.c_float,
                                     C.c_float, _fp]),
    'sgr_frames_to_uint8': (C.c_int, [_fp, _fp] + [C.c_int] * 5 + [_fp]),
    'sgr_packed_weight_bytes': (C.c_size_t, [C.c_int] * 5),
    'sgr_up_scratch_bytes': (C.c_size_t, [C.c_int] * 4),
    'sgr_pack_modconv_weight': (C.c_int, [_fp, _fp] + [C.c_int] * 7 + [_fp, _fp, _fp]),
    'sgr_choose_column_tile': (C.c_int, [C.c_int] * 4),
    'sgr_nchw_to_c8': (C.c_int, [_fp, _fp, _fp] + [C.c_int] * 6 + [_fp]),
    'sgr_modconv_forward': (C.c_int, [C.POINTER(ConvArgs), _fp]),
    'sgr_wgrad_scratch_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'sgr_modconv_wgrad': (C.c_int, [C.POINTER(WgradArgs), _fp]),
    'sgr_style_affine': (C.c_int, [_fp, C.c_int, C.c_int, _fp, _fp, C.c_int, _fp, _fp]),
    'sgr_demod': (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp]),
    'sgr_synthesis_workspace_bytes': (C.c_size_t, [C.POINTER(Synthesis), C.c_int]),
    'sgr_synthesis_backward_workspace_bytes': (C.c_size_t, [C.POINTER(Synthesis), C.c_int]),
    'sgr_synthesis_wgrad_scratch_bytes': (C.c_size_t, [C.POINTER(Synthesis), C.c_int]),
    'sgr_synthesis_backward': (C.c_int, [C.POINTER(Synthesis), _fp, C.c_int, C.POINTER(_fp), _fp, _fp, _fp, C.c_size_t,
                                         _fp]),
    'sgr_synthesis_backward_ex': (C.c_int, [C.POINTER(Synthesis), _fp, C.c_int, C.POINTER(_fp), _fp, _fp, _fp, C.c_size_t,
                                            C.POINTER(BackwardExtras), _fp]),
    'sgr_synthesis_forward': (C.c_int, [C.POINTER(Synthesis), _fp, C.c_int, _fp, _fp, C.c_size_t, C.POINTER(_fp),
                                        _fp]),
    'sgr_synthesis_forward_ex': (C.c_int, [C.POINTER(Synthesis), _fp, C.c_int, _fp, _fp, C.c_size_t, C.POINTER(_fp),
                                           C.POINTER(ForwardExtras), _fp]),
}

_lib = None


def lib():
    """Load libsgr.so once.  Raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libsgr.so is not built (%s); run `make -C %s` — there is no CPU fallback'
                               % (LIB_PATH, os.path.join(_HERE, 'csrc')))
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed: %s' % (what, lib().sgr_last_error().decode()))


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('libsgr needs CUDA tensors (got %s): there is no CPU fallback' % t.device)
    return t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
