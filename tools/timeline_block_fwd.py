"""Per-phase clock64 timeline of CTA 0 of block_fwd_umma at BASELINE size: python tools/timeline_block_fwd.py"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import torch
from wavenet import _lib
lib = _lib.load()
B, T, Cc, L = 1, 100000, 32, 50
M = B * T
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())
g = torch.Generator(device='cuda').manual_seed(0)
rnd = lambda *s: torch.randn(*s, device='cuda', generator=g)
x = rnd(M, Cc) * 0.5
xo = torch.empty(M, Cc, device='cuda')
zc = torch.empty(M, L * Cc, device='cuda')
wf, wg, wd = rnd(2, Cc, Cc) * 0.2, rnd(2, Cc, Cc) * 0.2, rnd(Cc, Cc) * 0.2
pb, bd = rnd(B, 2 * Cc) * 0.1, rnd(Cc) * 0.1
tl = torch.zeros(48, dtype=torch.int64, device='cuda')
names = ['loop top', 'tma landed', 'lo parts + MMA1 issued', 'MMA1 done', 'z computed + stored', 'sync before MMA2',
         'MMA2 done', 'tile end']
for d in (1, 512):
    for rep in range(3):
        lib.wn_block_fwd(p(x), p(xo), p(zc), L * Cc, p(wf), p(wg), p(wd), p(pb), p(bd), B, T, d, Cc, 0, st())
    lib.wn_debug_timeline(p(tl))
    lib.wn_block_fwd(p(x), p(xo), p(zc), L * Cc, p(wf), p(wg), p(wd), p(pb), p(bd), B, T, d, Cc, 0, st())
    torch.cuda.synchronize()
    lib.wn_debug_timeline(None)
    g = tl.cpu().numpy()[32:41].reshape(3, 3)
    g0 = g[:, 0].min()
    for k, nm in enumerate(['first CTA', 'middle CTA', 'last CTA']):
        print('  %-10s entry +%6d ns   first tile done +%6d ns   exit +%6d ns' % (nm, g[k, 0] - g0, g[k, 1] - g0, g[k, 2] - g0))
    t = tl.cpu().numpy()[:32].reshape(4, 8)
    print('d=%d  (cycles since first stamp; delta to previous phase)' % d)
    base = t[0, 0]
    prev = base
    for it in range(1):
        for i in range(8):
            if t[it, i] == 0:
                continue
            print('  tile %d %-26s %8d  (+%d)' % (it, names[i], t[it, i] - base, t[it, i] - prev))
            prev = t[it, i]
