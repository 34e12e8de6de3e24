"""Host-side mirror of wavenet/ops.py of the reference (same names, argument meaning and
error behaviour), calling the sm_100a kernels through the C ABI.

Reference symbols mirrored: optimizer_factory (ops.py:22-24), time_to_batch (:27-34),
batch_to_time (:37-43), causal_conv (:46-62), mu_law_encode (:65-73), mu_law_decode (:76-85).
"""
from __future__ import division

import math

import numpy as np
import torch

from . import _lib

# --------------------------------------------------------------------------------------
# tensor plumbing
# --------------------------------------------------------------------------------------


def _device():
    _lib.require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


def as_cuda(x, dtype):
    """numpy / python / torch -> contiguous CUDA tensor of `dtype`."""
    dev = _device()
    if isinstance(x, torch.Tensor):
        t = x.to(device=dev, dtype=dtype).contiguous()
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(x)), device=dev).to(dtype).contiguous()
    # the vectorised kernels want 16-byte aligned buffers: an offset view such as audio[1:] is contiguous
    # but not aligned -- give it its own allocation instead of failing with "bad argument"
    if t.numel() and t.data_ptr() % 16:
        t = t.clone()
    return t


# --------------------------------------------------------------------------------------
# mu-law tables (host logic): float32 restatement of ops.py:65-85 used ONLY to build the
# Q-1 decision thresholds / Q decode levels that make the integer kernels bit exact.
# --------------------------------------------------------------------------------------
_TABLES = {}


def _encode_formula_f32(audio, quantization_channels):
    f32 = np.float32
    mu = f32(quantization_channels - 1)
    audio = np.asarray(audio, dtype=np.float32)
    magnitude = (np.log(f32(1) + mu * np.abs(audio)).astype(np.float32) /
                 np.log(f32(1.0) + mu).astype(np.float32)).astype(np.float32)
    signal = (np.sign(audio).astype(np.float32) * magnitude).astype(np.float32)
    return ((signal + f32(1)) / f32(2) * mu + f32(0.5)).astype(np.float32).astype(np.int32)


def _decode_formula_f32(ids, quantization_channels):
    f32 = np.float32
    mu = quantization_channels - 1
    casted = np.asarray(ids).astype(np.float32)
    signal = (f32(2) * (casted / f32(mu)) - f32(1)).astype(np.float32)
    magnitude = (f32(1.0 / mu) * (np.power(f32(1 + mu), np.abs(signal)).astype(np.float32) - f32(1)))
    return (np.sign(signal).astype(np.float32) * magnitude.astype(np.float32)).astype(np.float32)


def _ord_to_f32(o):
    b = np.where(o < 0, (-o) | 0x80000000, o).astype(np.uint32)
    return b.view(np.float32)


def mu_law_tables(quantization_channels):
    """(thresholds[Q-1], lut[Q]) as float32 numpy arrays.

    thresholds[k-1] is the smallest float32 x in [-1, 1] whose encoding is >= k, found by
    bisection over the (monotone) float32 ordering of x."""
    q = int(quantization_channels)
    if q not in _TABLES:
        if q < 2:
            raise ValueError('quantization_channels must be >= 2, got {}'.format(q))
        one = np.int64(np.float32(1.0).view(np.int32))
        ks = np.arange(1, q, dtype=np.int32)
        lo = np.full(q - 1, -one, dtype=np.int64)   # encode(-1) = 0 < k
        hi = np.full(q - 1, one, dtype=np.int64)    # encode(+1) = Q-1 >= k
        while np.any(hi - lo > 1):
            mid = (lo + hi) // 2
            ge = _encode_formula_f32(_ord_to_f32(mid), q) >= ks
            hi = np.where(ge, mid, hi)
            lo = np.where(ge, lo, mid)
        thr = _ord_to_f32(hi).astype(np.float32)
        lut = _decode_formula_f32(np.arange(q), q)
        _TABLES[q] = (thr, lut)
    return _TABLES[q]


_DEV_TABLES = {}


def device_tables(quantization_channels):
    key = (int(quantization_channels), torch.cuda.current_device())
    if key not in _DEV_TABLES:
        thr, lut = mu_law_tables(quantization_channels)
        dev = _device()
        _DEV_TABLES[key] = (torch.as_tensor(thr, device=dev), torch.as_tensor(lut, device=dev))
    return _DEV_TABLES[key]


def mu_law_encode(audio, quantization_channels):
    '''Quantizes waveform amplitudes (ops.py:65-73).  Returns an int32 CUDA tensor of the
    input's shape.'''
    lib = _lib.load()
    x = as_cuda(audio, torch.float32)
    thr, _ = device_tables(quantization_channels)
    out = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    _lib.check(lib.wn_mulaw_encode(_lib.ptr(x), x.numel(), _lib.ptr(thr), int(quantization_channels),
                                   _lib.ptr(out), _lib.stream_ptr()), 'wn_mulaw_encode')
    return out


def mu_law_decode(output, quantization_channels):
    '''Recovers waveform from quantized values (ops.py:76-85).  float32 CUDA tensor.'''
    lib = _lib.load()
    ids = as_cuda(output, torch.int32)
    _, lut = device_tables(quantization_channels)
    out = torch.empty(ids.shape, dtype=torch.float32, device=ids.device)
    _lib.check(lib.wn_mulaw_decode(_lib.ptr(ids), ids.numel(), _lib.ptr(lut), int(quantization_channels),
                                   _lib.ptr(out), _lib.stream_ptr()), 'wn_mulaw_decode')
    return out


# --------------------------------------------------------------------------------------
# time <-> batch and the dilated causal convolution
# --------------------------------------------------------------------------------------


def time_to_batch(value, dilation, name=None):
    """ops.py:27-34.  Kept for API compatibility: the fused kernels never materialise it."""
    value = as_cuda(value, torch.float32)
    b, t, c = value.shape
    pad_elements = dilation - 1 - (t + dilation - 1) % dilation
    padded = torch.nn.functional.pad(value, (0, 0, 0, pad_elements))
    reshaped = padded.reshape(-1, dilation, c)
    transposed = reshaped.permute(1, 0, 2)
    return transposed.reshape(b * dilation, -1, c)


def batch_to_time(value, dilation, name=None):
    """ops.py:37-43."""
    value = as_cuda(value, torch.float32)
    s0, _, c = value.shape
    prepared = value.reshape(dilation, -1, c)
    transposed = prepared.permute(1, 0, 2)
    return transposed.reshape(s0 // dilation, -1, c).contiguous()


def causal_conv(value, filter_, dilation, name='causal_conv'):
    """ops.py:46-62: value [B,T,Cin], filter_ [W,Cin,Cout] -> [B,T,Cout]; dilation is addressing,
    no reshapes."""
    lib = _lib.load()
    x = as_cuda(value, torch.float32)
    w = as_cuda(filter_, torch.float32)
    if x.dim() != 3 or w.dim() != 3 or w.shape[1] != x.shape[2]:
        raise ValueError('causal_conv expects value [B,T,Cin] and filter [W,Cin,Cout]')
    b, t, cin = x.shape
    width, _, cout = w.shape
    y = torch.empty((b, t, cout), dtype=torch.float32, device=x.device)
    _lib.check(lib.wn_causal_conv(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), b, t, cin, cout, width,
                                  int(dilation), _lib.stream_ptr()), 'wn_causal_conv')
    return y


# --------------------------------------------------------------------------------------
# optimizers (TF-0.10 update rules, ops.py:6-24)
# --------------------------------------------------------------------------------------


class _Optimizer(object):
    """Holds hyper-parameters and slot buffers; `minimize(loss)` applies one update to the flat
    parameter buffer of the model that produced `loss` (the gradients were computed by the
    fused loss kernel sequence)."""

    def __init__(self, kind, learning_rate, momentum):
        self.kind = kind
        self.learning_rate = float(learning_rate)
        self.momentum = float(momentum)
        self.step = 0
        self._slots = None

    def _ensure_slots(self, params):
        if self._slots is None:
            if self.kind == 'adam':
                self._slots = (torch.zeros_like(params), torch.zeros_like(params))
            elif self.kind == 'sgd':
                self._slots = (torch.zeros_like(params),)
            else:  # rmsprop: `ms` starts at one (TF-0.10 RMSPropOptimizer), `mom` at zero
                self._slots = (torch.ones_like(params), torch.zeros_like(params))
        return self._slots

    def apply(self, params, grads, l2=0.0, grad_scale=1.0):
        lib = _lib.load()
        slots = self._ensure_slots(params)
        self.step += 1
        n, st = params.numel(), _lib.stream_ptr()
        if self.kind == 'adam':
            rc = lib.wn_optim_adam(_lib.ptr(params), _lib.ptr(grads), _lib.ptr(slots[0]), _lib.ptr(slots[1]), n,
                                   self.learning_rate, 0.9, 0.999, 1e-4, self.step, l2, grad_scale, st)
        elif self.kind == 'sgd':
            rc = lib.wn_optim_momentum(_lib.ptr(params), _lib.ptr(grads), _lib.ptr(slots[0]), n,
                                       self.learning_rate, self.momentum, l2, grad_scale, st)
        else:
            rc = lib.wn_optim_rmsprop(_lib.ptr(params), _lib.ptr(grads), _lib.ptr(slots[0]), _lib.ptr(slots[1]), n,
                                      self.learning_rate, 0.9, self.momentum, 1e-5, l2, grad_scale, st)
        _lib.check(rc, 'wn_optim_' + self.kind)

    def state_dict(self):
        """Step count and slot buffers (flat, in the layout of the model's flat parameter buffer) for checkpoints."""
        out = {'step': np.int64(self.step), 'kind': np.array(self.kind)}
        for i, t in enumerate(self._slots or ()):
            out['slot{}'.format(i)] = t.detach().cpu().numpy()
        return out

    def load_state_dict(self, state):
        if str(np.asarray(state.get('kind', self.kind))) != self.kind:
            raise ValueError('optimizer kind mismatch: checkpoint has {}, this is {}'.format(state['kind'], self.kind))
        self.step = int(state['step'])
        slots = [state[k] for k in sorted(k for k in state if k.startswith('slot'))]
        if slots:
            self._slots = tuple(as_cuda(a, torch.float32) for a in slots)

    def minimize(self, loss, var_list=None):
        """tf.train.Optimizer.minimize: `loss` is the tensor returned by WaveNetModel.loss()."""
        net = getattr(loss, '_wavenet_model', None)
        if net is None:
            raise ValueError('minimize() expects the tensor returned by WaveNetModel.loss()')
        self.apply(net.flat_params, net.flat_grads, l2=getattr(loss, '_wavenet_l2', 0.0))
        return loss


def create_adam_optimizer(learning_rate, momentum):
    return _Optimizer('adam', learning_rate, momentum)


def create_sgd_optimizer(learning_rate, momentum):
    return _Optimizer('sgd', learning_rate, momentum)


def create_rmsprop_optimizer(learning_rate, momentum):
    return _Optimizer('rmsprop', learning_rate, momentum)


optimizer_factory = {'adam': create_adam_optimizer,
                     'sgd': create_sgd_optimizer,
                     'rmsprop': create_rmsprop_optimizer}
