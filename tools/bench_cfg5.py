"""BASELINE config 5 (scaled net: R = D = 128, S = 512, 4 x dilations 1..512, T = 65536) training step timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 65536
kw = dict(batch_size=B, dilations=[2 ** i for i in range(10)] * 4, filter_width=2, residual_channels=128, dilation_channels=128,
          quantization_channels=256, skip_channels=512, use_biases=True)
net = wavenet.WaveNetModel(**kw, seed=0)
opt = wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9)
step = wavenet.TrainStep(net, opt, B, T)
rng = np.random.default_rng(0)
tt = np.arange(T) / 16000.0
a = np.clip(0.3 * np.sin(2 * np.pi * 220 * tt)[None] + 0.05 * rng.standard_normal((B, T)), -1, 1).astype(np.float32)
step.audio.copy_(torch.as_tensor(a))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('cfg5 (R=D=128, L=40, S=512) B=%d T=%d: %.2f ms/step, %.3e samples/s, loss %.4f' % (B, T, ms, B * T / ms * 1e3, float(step.loss)))
# per-stage times: events recorded after every launch of an eager step (serialises the side streams)
import ctypes as C
from wavenet import _lib
lib = _lib.load()
n_tags = 23
ms_tag, n_tag = (C.c_float * n_tags)(), (C.c_int32 * n_tags)()
torch.cuda.synchronize()
assert lib.wn_profile_begin() == 0
step._launch()
assert lib.wn_profile_end(ms_tag, n_tag, n_tags) == 0
buf = C.create_string_buffer(64)
tot = 0.0
for i in range(n_tags):
    lib.wn_profile_tag_name(i, buf, 64)
    if n_tag[i]:
        print('  %-18s %4d launches %8.3f ms' % (buf.value.decode(), n_tag[i], ms_tag[i]))
        tot += ms_tag[i]
print('  total %.3f ms (eager, serialised)' % tot)
