"""Where CTA 0 of the persistent forward kernel (block_fwd_chain) spends its cycles, per role."""
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
t = tl.cpu().numpy()
print('grid %d, CTA 0: %d tiles, %d cycles total (%.0f per tile)' % (t[8], t[7], t[0], t[0] / max(1, t[7])))
for k, nm in enumerate(['late flag waits + load issue', 'weight image wait', 'first product + x rows read, tile landed', 'z staged wait (flag polling, early loads)', 'dense done wait (layer change)', 'dense product issue']):
    print('  issuer    %-32s %9d  (%.0f per tile)' % (nm, t[1 + k], t[1 + k] / max(1, t[7])))
for k, nm in enumerate(['z staged wait', 'deferred publish', "x' staged wait", 'stores read wait', 'stores complete + publish']):
    print('  publisher %-32s %9d  (%.0f per tile)' % (nm, t[10 + k], t[10 + k] / max(1, t[7])))
