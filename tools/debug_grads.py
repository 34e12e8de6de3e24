import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
from wavenet import _lib
lib = _lib.load()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw = dict(batch_size=B, dilations=[2 ** i for i in range(10)] * 5, filter_width=2, residual_channels=32, dilation_channels=32,
          quantization_channels=256, skip_channels=512, use_biases=True)
rng = np.random.default_rng(0)
tt = np.arange(T) / 16000.0
a = np.clip(0.3 * np.sin(2 * np.pi * 220 * tt)[None] + 0.1 * rng.standard_normal((B, T)), -1, 1).astype(np.float32)
net = wavenet.WaveNetModel(**kw, seed=0)
l1 = float(net.loss(a)); g1 = net.gradients()
l1b = float(net.loss(a)); g1b = net.gradients()
lib.wn_debug_set_impl(0, 1)
net2 = wavenet.WaveNetModel(**kw, seed=0)
l2 = float(net2.loss(a)); g2 = net2.gradients()
lib.wn_debug_set_impl(0, 0)
print('loss umma %.6f rerun %.6f mma %.6f' % (l1, l1b, l2))
def re(x, y): return float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-30))
bad = [(re(g1[k], g2[k]), re(g1[k], g1b[k]), k) for k in g1]
bad.sort(reverse=True)
for e, e2, k in bad[:14]:
    print('%.3e (rerun %.1e)  %s' % (e, e2, k))
