"""One eager training step (loss + gradients) of BASELINE config 5 (R = D = 128, L = 40, T = 65536): the command the ncu captures run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, wavenet
net = wavenet.WaveNetModel(batch_size=1, dilations=[2 ** i for i in range(10)] * 4, filter_width=2, residual_channels=128,
                           dilation_channels=128, quantization_channels=256, skip_channels=512, use_biases=True, seed=0)
a = np.random.default_rng(0).uniform(-1, 1, (1, 65536)).astype(np.float32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    print(float(net.loss(a)))
