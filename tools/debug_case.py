import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
def l2(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))
B, T, L = 3, 4096, 30
rs = np.random.RandomState(5)
dil = [int(2 ** rs.randint(0, 10)) for _ in range(L)]
kw = dict(batch_size=B, dilations=dil, filter_width=2, residual_channels=32, dilation_channels=32,
          quantization_channels=256, skip_channels=64, use_biases=True)
audio = np.clip(0.4 * np.sin(np.arange(T) * 0.05)[None] + 0.2 * rs.randn(B, T), -1, 1).astype(np.float32)
out = sys.argv[1]
net = wavenet.WaveNetModel(**kw, seed=1)
gs = []
for _ in range(4):
    float(net.loss(audio)); gs.append(net.gradients())
np.savez(out, **{'%d|%s' % (i, k): v for i, g in enumerate(gs) for k, v in g.items()})
if len(sys.argv) > 2:
    ref = np.load(sys.argv[2])
    for i in range(4):
        worst = sorted(((l2(gs[i][k], ref['0|' + k]), k) for k in gs[i] if np.abs(ref['0|' + k]).max() > 0), reverse=True)[:4]
        print('run %d vs serialized:' % i, ' | '.join('%.1e %s' % (e, k.replace('wavenet/', '').replace('dilated_stack/', '')) for e, k in worst), flush=True)
