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
print('grid %d, CTA 0: %d items (item = DX of a layer + PRE of the layer below, one tile), %d cycles total (%.0f per item)' % (t[10], t[9], t[0], t[0] / n))
print('  loader     waits: flags not set %.0f (%d items), dx store left Ob %.0f, epilogue started item %.0f, DX MMAs done %.0f, PRE MMAs done %.0f per item' % (t[3] / n, t[21], t[4] / n, t[5] / n, t[6] / n, t[8] / n))
print('  MMA issuer waits: dpre tiles landed %.0f, x tiles landed %.0f, dx staged %.0f, weight image %.0f; MMA issue %.0f per item' % (t[1] / n, t[2] / n, t[22] / n, t[7] / n, t[20] / n))
print('epilogue thread 0: %d cycles total (%.0f per item)' % (t[11], t[11] / n))
print("  epilogue waits: dx' landed / next item (idle) %.0f, DX MMAs %.0f, dz landed %.0f, PRE MMAs %.0f, previous dpre store read %.0f per item" % (t[12] / n, t[13] / n, t[14] / n, t[15] / n, t[23] / n))
