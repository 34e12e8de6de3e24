"""Where CTA 0's issuer of the persistent backward kernel (block_bwd_chain) spends its cycles."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
from wavenet import _lib
lib = _lib.load()
p = json.load(open(os.path.join(ROOT, 'tensorflow-wavenet_b200', 'wavenet_params.json')))
net = wavenet.WaveNetModel(batch_size=1, dilations=p['dilations'], filter_width=2, residual_channels=32, dilation_channels=32,
                           quantization_channels=256, skip_channels=512, use_biases=True, seed=0)
a = np.random.default_rng(0).uniform(-1, 1, (1, 100000)).astype(np.float32)
float(net.loss(a))
tl = torch.zeros(48, dtype=torch.int64, device='cuda')
lib.wn_debug_timeline(C.c_void_p(tl.data_ptr()))
float(net.loss(a))
lib.wn_debug_timeline(None)
t = tl.cpu().numpy()[16:]
n = max(1, t[9])
print('grid %d, CTA 0: %d tiles, %d cycles total (%.0f per tile)' % (t[10], t[9], t[0], t[0] / n))
for k, nm in enumerate(['x tiles landed', "dz / dx' landed", 'neighbour dpre flags', 'dpre staged (epilogue 1)', 'shifted dpre rows landed',
                        'dx products done', 'dx store read (staging free)', 'weight image']):
    print('  issuer wait %-32s %9d  (%.0f per tile)' % (nm, t[1 + k], t[1 + k] / n))
