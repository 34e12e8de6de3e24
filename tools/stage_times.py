"""Real incremental cost of the stages of one training step: graph replay time of launch-sequence prefixes
(WN_TRUNCATE, api.cu).  Per-kernel event timing serialises launches that overlap in the real graph."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))
import numpy as np, torch, wavenet
p = json.load(open(os.path.join(ROOT, 'tensorflow-wavenet_b200', 'wavenet_params.json')))
B, T = 1, 100000
names = {1: 'forward layers', 2: '+ forward GEMMs + loss', 3: '+ post-processing gradient GEMMs', 0: '+ layer backward (full step)'}
prev = 0.0
for k in (1, 2, 3, 0):
    os.environ['WN_TRUNCATE'] = str(k)
    net = wavenet.WaveNetModel(batch_size=B, dilations=p['dilations'], filter_width=2, residual_channels=32, dilation_channels=32,
                               quantization_channels=256, skip_channels=512, use_biases=True, seed=0)
    opt = wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9)
    step = wavenet.TrainStep(net, opt, B, T, use_cuda_graph=True)
    a = torch.tensor(np.random.default_rng(0).uniform(-1, 1, (B, T)).astype(np.float32), device='cuda')
    step.audio.copy_(a)
    for _ in range(5):
        step._graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for _ in range(n):
        step._graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('%-36s %.3f ms  (+%.3f)' % (names[k], ms, ms - prev))
    prev = ms
    del step, net
