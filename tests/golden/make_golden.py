"""Writes tests/golden/oracle_small.npz.

The reference cannot be imported here (it needs TensorFlow 0.10 + librosa; neither is
installable), so the fixture holds (a) the reference's own known-answer vectors for mu-law
(test/test_mu_law.py:113-124 and the seed-42 input of :126-137, encoded by the oracle, which
test_oracle_golden.py pins to the reference's float32 numpy formulas) and (b) oracle outputs
for one small seeded network, so that the GPU parity tests can also run against frozen numbers.
Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import wavenet_oracle as O  # noqa: E402

kw = dict(batch_size=2, dilations=[1, 2, 4, 8, 16, 32] * 2, filter_width=2, residual_channels=32,
          dilation_channels=32, quantization_channels=256, skip_channels=64, use_biases=True,
          global_condition_channels=4, global_condition_cardinality=3)
seed = 5
net = O.OracleWaveNet(seed=seed, bias_scale=0.1, faithful=True, **kw)
rng = np.random.default_rng(seed)
t = np.arange(600) / 2000.0
audio = np.clip(0.4 * np.sin(2 * np.pi * 155.56 * t)[None] + 0.1 * rng.standard_normal((2, 600)), -1, 1)
audio = audio.astype(np.float32)
gc = np.array([2, 0])
loss, logits, grads = net.loss_and_grads(audio, gc)
np.random.seed(42)
mx = np.concatenate([np.array([-1.0, 1.0, 0.6, -0.25, 0.01, 0.33, -0.9999, 0.42, 0.1, -0.45], np.float32),
                     np.random.uniform(-1, 1, 2048).astype(np.float32)])
sd = net.state_dict()
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'oracle_small.npz'), config=json.dumps(kw), seed=seed,
                    audio=audio, gc=gc, loss=np.float64(loss), logits=logits,
                    grad_post2=grads['wavenet/postprocessing/postprocess2'],
                    grad_causal=grads['wavenet/causal_layer/filter'], mulaw_x=mx, mulaw_ids=O.mu_law_encode(mx, 256),
                    **{'w:' + k: v for k, v in sd.items()})
print('loss', loss, 'logits', logits.shape)
