import os, sys
os.environ['WN_DEBUG_SYNC'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dil = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 4]
net = wavenet.WaveNetModel(batch_size=1, dilations=dil, filter_width=2, residual_channels=32, dilation_channels=32,
                           quantization_channels=256, skip_channels=64, use_biases=True, seed=0)
a = np.random.default_rng(0).uniform(-1, 1, (1, T)).astype(np.float32)
try:
    print('dil', dil, 'loss', float(net.loss(a)))
except Exception as e:
    print('dil', dil, 'FAILED', repr(e)[:300])
