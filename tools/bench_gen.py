"""Fast-generation timing: python tools/bench_gen.py [streams] [n_samples]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
p = json.load(open(os.path.join(ROOT, 'tensorflow-wavenet_b200', 'wavenet_params.json')))
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
net = wavenet.WaveNetModel(batch_size=1, dilations=p['dilations'], filter_width=2, residual_channels=32, dilation_channels=32,
                           quantization_channels=256, skip_channels=512, use_biases=True, seed=0)
first = np.random.RandomState(0).randint(0, 256, streams)
u = np.random.RandomState(1).random_sample((streams, n))
net.generate(16, first, uniforms=u[:, :16])
torch.cuda.synchronize(); t0 = time.perf_counter()
net.generate(n, first, uniforms=u)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('streams %d  n %d  SPB=%s: %.2f us/step  %.0f samples/s/stream  %.3e aggregate' % (
    streams, n, os.environ.get('WN_GEN_SPB', 'auto'), dt / n * 1e6, n / dt, streams * n / dt))
