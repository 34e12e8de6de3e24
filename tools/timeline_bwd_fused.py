"""Where epilogue thread 0 of CTA 0 of the fused backward kernel (block_bwd_chain_f_kernel) spends its cycles."""
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
print('grid %d, CTA 0: %d items, %d cycles (%.0f per item)' % (t[10], n, t[0], t[0] / n))
print('  waits per item: tiles landed (idle) %.0f, DX MMAs %.0f, PRE MMAs %.0f' % (t[12] / n, t[13] / n, t[15] / n))
print('  flushes %d: %.0f cycles each; PRE part after a flush %.0f, otherwise %.0f cycles' %
      (t[25], t[24] / max(1, t[25]), t[26] / max(1, t[25]), t[27] / max(1, t[28])))
print('  loader waits per item: flags not set %.0f (%d items), epilogue inside PRE %.0f, DX+PRE MMAs read their tiles %.0f, dz MMA %.0f, weight-gradient MMAs %.0f, dx store left Ob %.0f'
      % (t[3] / n, t[21], t[5] / n, t[6] / n, t[8] / n, t[7] / n, t[4] / n))
