"""Wide (R, D multiples of 64) residual blocks in 16-bit storage vs the fp64 oracle: loss / logits / gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('tensorflow-wavenet_b200', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch, wavenet
import wavenet_oracle as O
from wn_helpers import make_pair, rel_err, l2_rel
from test_gpu_model import CASES, _audio
for case in sys.argv[1:] or ['wide_r64_d64', 'scaled_r128_short']:
    kw, T, gc = CASES[case]
    onet, net = make_pair(O, wavenet, seed=1, **kw)
    audio = _audio(np.random.default_rng(7), kw['batch_size'], T)
    loss_ref, logits_ref, grads_ref = onet.loss_and_grads(audio, gc)
    loss = float(net.loss(audio, gc))
    ids = O.mu_law_encode(audio, kw['quantization_channels'])
    logits = net.logits(ids, gc).cpu().numpy()
    got = net.gradients()
    print(case, 'loss', loss, loss_ref, 'rel', abs(loss - loss_ref) / abs(loss_ref), 'logits rel', rel_err(logits, logits_ref))
    worst = []
    for k, g in grads_ref.items():
        if np.abs(g).max() == 0: continue
        cos = float(np.dot(got[k].ravel().astype(np.float64), g.ravel().astype(np.float64)) / max(np.linalg.norm(got[k]) * np.linalg.norm(g), 1e-300))
        worst.append((l2_rel(got[k], g), cos, k))
    worst.sort(reverse=True)
    for w in worst[:6]: print('   %.3e cos %.5f %s' % w)
