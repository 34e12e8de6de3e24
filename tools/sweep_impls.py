"""Randomised cross-check of the tcgen05 path against the mma.sync path of the same arithmetic (loss, logits, every
gradient, run-to-run reproducibility) over odd shapes: python tools/sweep_impls.py [n_cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
from wavenet import _lib
lib = _lib.load()
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
def l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))
bad = 0
for case in range(n_cases):
    B = int(rs.choice([1, 1, 2, 3, 5]))
    T = int(rs.choice([1, 7, 127, 128, 129, 255, 300, 1000, 1337, 4096, 5000, 20011, 70001]))
    if B * T > 150000: B = 1
    L = int(rs.choice([1, 2, 3, 9, 14, 30]))
    dil = [int(2 ** rs.randint(0, 10)) for _ in range(L)]
    gc = rs.rand() < 0.4
    kw = dict(batch_size=B, dilations=dil, filter_width=2, residual_channels=32, dilation_channels=32,
              quantization_channels=int(rs.choice([256, 128, 64])), skip_channels=int(rs.choice([32, 64, 256, 512, 100])),
              use_biases=bool(rs.rand() < 0.7), residual_postproc=bool(rs.rand() < 0.2))
    if gc:
        kw.update(global_condition_channels=int(rs.choice([4, 16, 32])), global_condition_cardinality=7)
    ids_gc = rs.randint(0, 7, B) if gc else None
    audio = np.clip(0.4 * np.sin(np.arange(T) * 0.05)[None] + 0.2 * rs.randn(B, T), -1, 1).astype(np.float32)
    res = {}
    try:
        for name, flag in (('umma', 0), ('umma2', 0), ('mma', 1)):
            lib.wn_debug_set_impl(flag, flag)
            net = wavenet.WaveNetModel(**kw, seed=case)
            # non-zero biases so that they matter
            sd = net.state_dict()
            r2 = np.random.RandomState(case)
            for k in sd:
                if 'bias' in k: sd[k] = (0.1 * r2.randn(*sd[k].shape)).astype(np.float32)
            net.load_state_dict(sd)
            loss = float(net.loss(audio, ids_gc))
            res[name] = (loss, net.gradients())
    finally:
        lib.wn_debug_set_impl(0, 0)
    worst = max(l2(res['umma'][1][k], g) for k, g in res['mma'][1].items() if np.abs(g).max() > 0)
    rerun, rkey = max((l2(res['umma'][1][k], g), k) for k, g in res['umma2'][1].items() if np.abs(g).max() > 0)
    dl = abs(res['umma'][0] - res['mma'][0]) / max(abs(res['mma'][0]), 1e-30)
    ok = dl < 3e-4 and worst < 3e-2 and rerun < 1e-4 and np.isfinite(res['umma'][0])
    bad += 0 if ok else 1
    print('%s B=%d T=%-6d L=%-2d S=%-3d Q=%-3d gc=%d bias=%d rp=%d d=%s | loss rel %.1e grad l2 %.1e rerun %.1e' % (
        'ok  ' if ok else 'FAIL', B, T, L, kw['skip_channels'], kw['quantization_channels'], gc, kw['use_biases'],
        kw['residual_postproc'], dil[:6], dl, worst, rerun), rkey.replace('wavenet/', '') if not ok else '', flush=True)
print('%d / %d cases failed' % (bad, n_cases))
