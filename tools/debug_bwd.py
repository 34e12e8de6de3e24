"""Stand-alone residual block (wn_block_fwd / wn_block_bwd on the production kernels) against the fp64 oracle:
prints the max-norm relative error of every output for a list of shapes.  python tools/debug_bwd.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('tensorflow-wavenet_b200', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
from wavenet import _lib
from wn_helpers import dev, p, rel_err, stream
from test_gpu_kernels import _block_oracle

lib = _lib.load()
cases = [(1, 1000, 8, 0), (1, 1000, 64, 0), (1, 1000, 128, 0), (1, 1000, 1, 0), (1, 1000, 4, 0), (3, 333, 8, 0), (2, 700, 512, 0),
         (1, 999, 4, 1), (1, 16, 2, 0), (2, 1153, 127, 0)]
C_ = 32
for B, T, d, is_last in cases:
    rng = np.random.default_rng(T + d)
    M = B * T
    lim = np.sqrt(6.0 / (4 * C_))
    x = rng.standard_normal((B, T, C_)).astype(np.float32)
    wf = rng.uniform(-lim, lim, (2, C_, C_)).astype(np.float32)
    wg = rng.uniform(-lim, lim, (2, C_, C_)).astype(np.float32)
    wd = rng.uniform(-lim, lim, (1, C_, C_)).astype(np.float32)
    prebias = (0.2 * rng.standard_normal((B, 2 * C_))).astype(np.float32)
    bd = (0.2 * rng.standard_normal(C_)).astype(np.float32)
    gz = rng.standard_normal((B, T, C_)).astype(np.float32)
    gx = rng.standard_normal((B, T, C_)).astype(np.float32)
    ref = _block_oracle(x, wf, wg, wd, prebias, bd, d, is_last, gz, gx)
    ldz = 3 * C_
    dx_, dwf_, dwg_, dwd_, dpb_, dbd_ = dev(x), dev(wf), dev(wg), dev(wd), dev(prebias), dev(bd)
    zc = torch.zeros(M, ldz, device='cuda')
    xo = torch.zeros(M, C_, device='cuda')
    rc = lib.wn_block_fwd(p(dx_), p(xo), C.c_void_p(zc.data_ptr() + 4 * C_), ldz, p(dwf_), p(dwg_), p(dwd_),
                          p(dpb_), p(dbd_), B, T, d, C_, is_last, stream())
    torch.cuda.synchronize()
    z = zc[:, C_:2 * C_].cpu().numpy().reshape(B, T, C_)
    out = {'z': rel_err(z, ref['z'])}
    dzs = torch.zeros(M, ldz, device='cuda')
    dzs[:, C_:2 * C_] = dev(gz).reshape(M, C_)
    dxo = dev(gx).reshape(M, C_)
    dx = torch.zeros(M, C_, device='cuda')
    dpre = torch.zeros(M, 2 * C_, device='cuda')
    gwf, gwg, gwd = torch.zeros(2, C_, C_, device='cuda'), torch.zeros(2, C_, C_, device='cuda'), torch.zeros(C_, C_, device='cuda')
    gpb, gbd = torch.zeros(B, 2 * C_, device='cuda'), torch.zeros(C_, device='cuda')
    rc2 = lib.wn_block_bwd(p(dx_), p(dxo), C.c_void_p(dzs.data_ptr() + 4 * C_), ldz, p(dx), p(dpre),
                           C.c_void_p(zc.data_ptr() + 4 * C_), p(dwf_), p(dwg_), p(dwd_), p(dpb_), p(gwf), p(gwg),
                           p(gwd), p(gpb), p(gbd), B, T, d, C_, is_last, stream())
    torch.cuda.synchronize()
    dxh = dx.cpu().numpy().reshape(B, T, C_)
    out['dx'] = rel_err(dxh, ref['dx'])
    # where is the dx error?  per-row max error, first / last bad rows
    err_rows = np.abs(dxh - ref['dx']).max(axis=2) / np.abs(ref['dx']).max()
    bad = np.argwhere(err_rows > 4e-3)
    out['gwf0'] = rel_err(gwf[0].cpu().numpy(), ref['dwf'][0]); out['gwf1'] = rel_err(gwf[1].cpu().numpy(), ref['dwf'][1])
    out['gwg0'] = rel_err(gwg[0].cpu().numpy(), ref['dwg'][0]); out['gwg1'] = rel_err(gwg[1].cpu().numpy(), ref['dwg'][1])
    out['gpb'] = rel_err(gpb.cpu().numpy(), ref['dpb'])
    if not is_last:
        out['gwd'] = rel_err(gwd.cpu().numpy(), ref['dwd'][0]); out['gbd'] = rel_err(gbd.cpu().numpy(), ref['dbd'])
    print('B{} T{} d{} last{} rc {} {}: '.format(B, T, d, is_last, rc, rc2) + ' '.join('{} {:.1e}'.format(k, v) for k, v in out.items()),
          '| bad dx rows: {} first {} last {}'.format(len(bad), bad[:3].tolist(), bad[-2:].tolist()))
