"""clock64 phases of CTA 0 of a block_fwd_h launch (a d == 32 layer) inside one eager training step."""
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
t = tl.cpu().numpy()[:32].reshape(4, 8)
names = ['loop top', 'tma landed', 'x row read, MMA1 issued', 'MMA1 done', 'z computed + staged', 'after sync 1',
         'MMA2 done', 'tile end (after sync 2)']
base = prev = t[0, 0]
for it in range(4):
    for i in range(8):
        if t[it, i] == 0:
            continue
        print('tile %d %-26s %8d  (+%d)' % (it, names[i], t[it, i] - base, t[it, i] - prev))
        prev = t[it, i]
