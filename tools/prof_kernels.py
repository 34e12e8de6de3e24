"""Launches the hot kernels a few times at BASELINE size (for ncu / timing): python tools/prof_kernels.py [which]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np
import torch
from wavenet import _lib

lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, T, Cc, L = 1, 100000, 32, 50
M = B * T
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())
g = torch.Generator(device='cuda').manual_seed(0)
rnd = lambda *s: torch.randn(*s, device='cuda', generator=g)


def timeit(name, fn, bytes_=None, flops=None):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    extra = ''
    if bytes_:
        extra += ' %.0f GB/s' % (bytes_ / us / 1e3)
    if flops:
        extra += ' %.0f TFLOP/s' % (flops / us / 1e6)
    print('%-28s %9.1f us%s' % (name, us, extra), flush=True)


if which in ('all', 'block_fwd'):
    x = rnd(M, Cc) * 0.5
    xo = torch.empty(M, Cc, device='cuda')
    zc = torch.empty(M, L * Cc, device='cuda')
    zct = torch.empty(L * Cc, M, device='cuda')
    wf, wg, wd = rnd(2, Cc, Cc) * 0.2, rnd(2, Cc, Cc) * 0.2, rnd(Cc, Cc) * 0.2
    pb, bd = rnd(B, 2 * Cc) * 0.1, rnd(Cc) * 0.1
    for d in (1, 64, 512):
        timeit('block_fwd d=%d' % d, lambda: lib.wn_block_fwd(p(x), p(xo), p(zc), L * Cc, p(wf), p(wg), p(wd), p(pb), p(bd),
                                                               B, T, d, Cc, 0, st()), bytes_=M * 384.0)
if which in ('all', 'block_bwd'):
    x = rnd(M, Cc) * 0.5
    dxo = rnd(M, Cc) * 0.1
    dzs = rnd(M, L * Cc) * 0.1
    zc = rnd(M, L * Cc) * 0.3
    dx = torch.empty(M, Cc, device='cuda')
    dpre = torch.empty(M, 2 * Cc, device='cuda')
    wf, wg, wd = rnd(2, Cc, Cc) * 0.2, rnd(2, Cc, Cc) * 0.2, rnd(Cc, Cc) * 0.2
    pb = rnd(B, 2 * Cc) * 0.1
    gwf, gwg, gwd = torch.zeros(2, Cc, Cc, device='cuda'), torch.zeros(2, Cc, Cc, device='cuda'), torch.zeros(Cc, Cc, device='cuda')
    gpb, gbd = torch.zeros(B, 2 * Cc, device='cuda'), torch.zeros(Cc, device='cuda')
    for d in (1, 512):
        timeit('block_bwd d=%d' % d, lambda: lib.wn_block_bwd(p(x), p(dxo), p(dzs), L * Cc, p(dx), p(dpre), p(zc), p(wf), p(wg),
                                                               p(wd), p(pb), p(gwf), p(gwg), p(gwd), p(gpb), p(gbd), B, T, d,
                                                               Cc, 0, st()), bytes_=M * 640.0)
if which in ('all', 'gemm'):
    S, Q, LD = 512, 256, L * Cc
    shapes = [('skip_fwd', M, S, LD), ('post1_fwd', M, S, S), ('post2_fwd', M, Q, S), ('post2_dgrad', M, S, Q),
              ('skip_dgrad', M, LD, S), ('skip_wgrad', LD, S, M), ('post1_wgrad', S, S, M), ('post2_wgrad', S, Q, M)]
    for name, m, n, k in shapes:
        a = rnd(m, k)
        b = rnd(n, k)
        c = torch.zeros(m, n, device='cuda')
        split = 0
        flags = 0
        if name.endswith('wgrad'):
            tiles = ((m + 127) // 128) * ((n + 255) // 256)
            split = max(1, (4 * 148 + tiles - 1) // tiles)
            flags = 4
        timeit('gemm ' + name, lambda: lib.wn_gemm_nt_umma(p(a), k, p(b), k, p(c), n, None, 0, m, n, k, None, None, 0, flags,
                                                            split, st()), flops=2.0 * m * n * k)
print('done')
