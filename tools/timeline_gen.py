"""Hop-by-hop %globaltimer timeline of one sample of the latency-mode generator: python tools/timeline_gen.py"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
from wavenet import _lib
lib = _lib.load()
p = json.load(open(os.path.join(ROOT, 'tensorflow-wavenet_b200', 'wavenet_params.json')))
net = wavenet.WaveNetModel(batch_size=1, dilations=p['dilations'], filter_width=2, residual_channels=32, dilation_channels=32,
                           quantization_channels=256, skip_channels=512, use_biases=True, seed=0)
tl = torch.zeros(48, dtype=torch.int64, device='cuda')
net.generate(64, [128], seed=0)
lib.wn_debug_timeline(C.c_void_p(tl.data_ptr()))
n = 4000
torch.cuda.synchronize(); t0 = time.perf_counter()
net.generate(n, [128], seed=0)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
lib.wn_debug_timeline(None)
print('%.2f us/sample over %d samples (incl. launch)' % (dt / n * 1e6, n))
t = tl.cpu().numpy()
names = {0: 'head has id'}
for c in range(7): names[1 + c] = 'chain CTA %d done' % c
names.update({8: 'post0 saw zdone', 9: 'post0 v0 out', 10: 'post0 v1 out', 11: 'post0 logits out', 12: 'sampler softmax done', 13: 'sampler id out', 15: 'head has id (next step)'})
prev = t[0]
for k in sorted(names):
    if t[k]:
        print('  %-26s +%6d ns (+%d)' % (names[k], t[k] - t[0], t[k] - prev)); prev = t[k]
