"""One eager training step (loss + gradients) at the BASELINE size: the command the ncu captures in profiles/ run."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, wavenet
p = json.load(open(os.path.join(ROOT, 'tensorflow-wavenet_b200', 'wavenet_params.json')))
net = wavenet.WaveNetModel(batch_size=1, dilations=p['dilations'], filter_width=2, residual_channels=32, dilation_channels=32,
                           quantization_channels=256, skip_channels=512, use_biases=True, seed=0)
a = np.random.default_rng(0).uniform(-1, 1, (1, 100000)).astype(np.float32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    print(float(net.loss(a)))
