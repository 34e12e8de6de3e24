"""Pipelined generator: where does a step go?  Priming (forced ids: chain only) vs sampling, per stream count."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
p = json.load(open(os.path.join(ROOT, 'tensorflow-wavenet_b200', 'wavenet_params.json')))
net = wavenet.WaveNetModel(batch_size=1, dilations=p['dilations'], filter_width=2, residual_channels=32, dilation_channels=32,
                           quantization_channels=256, skip_channels=512, use_biases=True, seed=0)
n = 3000
for streams in (1, 2, 4, 8, 16, 32):
    ids = np.random.RandomState(0).randint(0, 256, (streams, n)).astype(np.int32)
    u = np.random.RandomState(1).random_sample((streams, n))
    net.prime(ids[:, :16])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    net.prime(ids)
    torch.cuda.synchronize(); tp = time.perf_counter() - t0
    net.generate(16, ids[:, 0], uniforms=u[:, :16])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    net.generate(n, ids[:, 0], uniforms=u)
    torch.cuda.synchronize(); tg = time.perf_counter() - t0
    print('streams %2d: priming %.2f us/step (%.2f per stream), sampling %.2f us/step (%.2f per stream, %.0f samples/s/stream)' % (
        streams, tp / n * 1e6, tp / n * 1e6 / streams, tg / n * 1e6, tg / n * 1e6 / streams, n / tg))
