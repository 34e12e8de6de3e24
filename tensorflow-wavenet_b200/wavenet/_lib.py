"""ctypes binding of libwavenet_b200.so (the C ABI declared in include/wavenet_b200.h).

PyTorch is used only to own device memory and streams; every argument that crosses this
boundary is a raw device pointer, a size or a plain C struct.  There is no CPU fallback:
importing this module without the built library raises, and so does any call on a machine
without a CUDA device.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'libwavenet_b200.so')
WN_MAX_LAYERS = 128


class WnConfig(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('residual_channels', C.c_int32),
                ('dilation_channels', C.c_int32), ('skip_channels', C.c_int32),
                ('quantization_channels', C.c_int32), ('gc_channels', C.c_int32),
                ('gc_cardinality', C.c_int32), ('use_biases', C.c_int32),
                ('residual_postproc', C.c_int32), ('dilations', C.c_int32 * WN_MAX_LAYERS),
                ('scalar_input', C.c_int32), ('initial_filter_width', C.c_int32)]


LAYOUT_FIELDS = ['causal', 'filter', 'gate', 'dense', 'skip', 'gc_filter', 'gc_gate', 'filter_bias',
                 'gate_bias', 'dense_bias', 'skip_bias', 'post1', 'post2', 'post1_bias', 'post2_bias',
                 'gc_embedding', 'total']      # (field order of the C struct, not the order of the groups in memory)


class WnLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in LAYOUT_FIELDS]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F = C.c_float
_D = C.c_double
_CFG = C.POINTER(WnConfig)

# name -> (restype, argtypes); must list every symbol include/wavenet_b200.h declares
SIGNATURES = {
    'wn_abi_version': (C.c_int, []),
    'wn_param_layout': (C.c_int, [_CFG, C.POINTER(WnLayout)]),
    'wn_mulaw_encode': (C.c_int, [_P, _I64, _P, _I32, _P, _P]),
    'wn_mulaw_decode': (C.c_int, [_P, _I64, _P, _I32, _P, _P]),
    'wn_frontend_fwd': (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    'wn_frontend_bwd': (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    'wn_causal_conv': (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    'wn_block_fwd': (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    'wn_block_bwd': (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               _I32, _I32, _I32, _I32, _I32, _P]),
    'wn_gemm_tf32': (C.c_int, [_I32, _P, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32,
                               _I32, _P]),
    'wn_gemm_nt_umma': (C.c_int, [_P, _I32, _P, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32,
                                  _I32, _P]),
    'wn_profile_mark': (C.c_int, [_I32, _P]),
    'wn_debug_timeline': (C.c_int, [_P]),
    'wn_debug_trap_info': (C.c_int, [_P]),
    'wn_debug_set_gen_impl': (C.c_int, [_I32]),
    'wn_gemm_umma': (C.c_int, [_I32, _P, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32,
                               _I32, _P]),
    'wn_gemm_f16_nt': (C.c_int, [_P, _I32, _P, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _I32, C.c_float,
                                 _I32, _P]),
    'wn_gemm_f16_tn': (C.c_int, [_P, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, C.c_float, _I32, _P]),
    'wn_softmax_xent': (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _I32, _P, _I32, _P]),
    'wn_gemm_f16_nt_colsum': (C.c_int, [_P, _I32, _P, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, _F, _P, _F, _P]),
    'wn_post2_xent': (C.c_int, [_P, _I32, _P, _I32, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _F, _P, _F, _P]),
    'wn_train_workspace_bytes': (_I64, [_CFG, _I32, _I32]),
    'wn_loss_grad': (C.c_int, [_CFG, _P, _P, _P, _I64, _P, _P, _P, _I32, _I32, _P, _P]),
    'wn_forward_workspace_bytes': (_I64, [_CFG, _I32, _I32]),
    'wn_forward_logits': (C.c_int, [_CFG, _P, _P, _I64, _P, _P, _I32, _I32, _P, _P]),
    'wn_optim_adam': (C.c_int, [_P, _P, _P, _P, _I64, _D, _D, _D, _D, _I64, _F, _F, _P]),
    'wn_optim_momentum': (C.c_int, [_P, _P, _P, _I64, _D, _D, _F, _F, _P]),
    'wn_optim_rmsprop': (C.c_int, [_P, _P, _P, _P, _I64, _D, _D, _D, _D, _F, _F, _P]),
    'wn_gen_state_bytes': (_I64, [_CFG, _I32]),
    'wn_gen_reset': (C.c_int, [_CFG, _P, _I32, _P]),
    'wn_gen_run': (C.c_int, [_CFG, _P, _P, _I32, _P, _P, _P, _P, _I32, _F, _I32, _P, _P, _P]),
    'wn_gen_commit': (C.c_int, [_CFG, _P, _I32, _P]),
    'wn_sample': (C.c_int, [_P, _P, _I32, _I32, _P, _P]),
    'wn_debug_set_impl': (C.c_int, [_I32, _I32]),
    'wn_set_grad_ready_event': (C.c_int, [_P]),
    'wn_predict_last': (C.c_int, [_CFG, _P, _P, _I64, _P, _P, _I32, _I32, _P, _P]),
    'wn_add_l2': (C.c_int, [_P, _P, _I64, _F, _P]),
    'wn_profile_begin': (C.c_int, []),
    'wn_profile_end': (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32), _I32]),
    'wn_profile_tag_name': (C.c_int, [_I32, C.c_char_p, _I32]),
}

_lib = None


def load():
    """dlopen the in-tree library (built by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError('libwavenet_b200.so is missing at {}: build it with '
                              '`make -C tensorflow-wavenet_b200/csrc` (there is no CPU fallback)'
                              .format(LIB_PATH))
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class WavenetCudaError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        kind = 'bad argument / unsupported shape' if rc < 0 else 'cudaError'
        raise WavenetCudaError('{} failed: {} {}'.format(what, kind, rc))


def require_cuda():
    if not torch.cuda.is_available():
        raise WavenetCudaError('wavenet_b200 needs a CUDA device (sm_100a); there is no CPU fallback')


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a (contiguous, CUDA) tensor, or NULL for None."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), 'expected a contiguous CUDA tensor'
    return C.c_void_p(t.data_ptr())


def make_config(dilations, residual_channels, dilation_channels, skip_channels, quantization_channels,
                gc_channels, gc_cardinality, use_biases, residual_postproc, scalar_input=False, initial_filter_width=32):
    if len(dilations) > WN_MAX_LAYERS:
        raise ValueError('at most {} layers are supported'.format(WN_MAX_LAYERS))
    cfg = WnConfig()
    cfg.n_layers = len(dilations)
    cfg.residual_channels = residual_channels
    cfg.dilation_channels = dilation_channels
    cfg.skip_channels = skip_channels
    cfg.quantization_channels = quantization_channels
    cfg.gc_channels = gc_channels or 0
    cfg.gc_cardinality = gc_cardinality or 0
    cfg.use_biases = 1 if use_biases else 0
    cfg.residual_postproc = 1 if residual_postproc else 0
    cfg.scalar_input = 1 if scalar_input else 0
    cfg.initial_filter_width = int(initial_filter_width)
    for i, d in enumerate(dilations):
        cfg.dilations[i] = int(d)
    return cfg


def param_layout(cfg):
    lo = WnLayout()
    rc = load().wn_param_layout(C.byref(cfg), C.byref(lo))
    if rc != 0:
        raise NotImplementedError(
            'no sm_100a kernel for this configuration (rc={}): this build needs residual_channels in '
            '{{4, 8, ..., 256}} (a divisor of 256), dilation / skip / quantization channels multiples of 4 and '
            'quantization_channels <= 1024'.format(rc))
    return lo
