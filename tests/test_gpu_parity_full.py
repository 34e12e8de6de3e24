"""Parity at the sizes and configurations BASELINE.json names (cfg2 .. cfg4), against the CPU oracle.

cfg2  default wavenet_params.json, B = 1 x T = 100000 (the benchmarked shape: 782 tiles per layer, every CTA of the
      persistent kernels busy): loss, logits and gradients against the fp32 oracle.
cfg3  default params + global conditioning (gc_channels = 32, gc_cardinality = 377).
cfg4  fast generation on default params: > 5117 (receptive field) teacher-forced steps through BOTH generator kernels so
      that every delay line has wrapped; 256 streams against the batched oracle; temperature != 1 against
      generate.py:229-241 (O.scale_prediction + O.choice_from_uniform).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import wavenet_oracle as O
from wn_helpers import GRAD_L2_VS_EXACT, LOGIT_RTOL, LOSS_RTOL, l2_rel, make_pair, matched_oracle, rel_err

DEFAULT_NET = dict(batch_size=1, dilations=[2 ** i for i in range(10)] * 5, filter_width=2, residual_channels=32,
                   dilation_channels=32, quantization_channels=256, skip_channels=512, use_biases=True)
TEST_NET = dict(batch_size=1, dilations=[1, 2, 4, 8, 16, 32, 64] * 2, filter_width=2, residual_channels=32,
                dilation_channels=32, quantization_channels=256, skip_channels=32)


def _audio(rng, b, t):
    tt = np.arange(t) / 16000.0
    a = (0.3 * np.sin(2 * np.pi * 220 * tt)[None] + 0.3 * np.sin(2 * np.pi * 331 * tt)[None] +
         0.1 * rng.standard_normal((b, t)))
    return np.clip(a, -1, 1).astype(np.float32)


# ----------------------------------------------------------------------------------------- cfg2
def test_cfg2_full_size_vs_oracle():
    """B = 1 x T = 100000, default params: loss 1e-3, logits 1e-3 (max-norm), gradients norm-wise against the fp32
    oracle and tightly against the oracle that emulates the kernels' 11-bit operand rounding."""
    import wavenet
    torch.set_num_threads(os.cpu_count() or 8)
    onet, net = make_pair(O, wavenet, seed=2, dtype=torch.float32, **DEFAULT_NET)
    audio = _audio(np.random.default_rng(11), 1, 100000)
    loss_ref, logits_ref, grads_ref = onet.loss_and_grads(audio)
    loss = float(net.loss(audio))
    got = net.gradients()
    assert abs(loss - loss_ref) <= LOSS_RTOL * abs(loss_ref), (loss, loss_ref)
    ids = O.mu_law_encode(audio, 256)
    logits = net.logits(ids).cpu().numpy()
    assert rel_err(logits, logits_ref) < LOGIT_RTOL
    del logits, logits_ref
    _, _, grads_m = matched_oracle(O, onet, **DEFAULT_NET).loss_and_grads(audio)
    keys = ['wavenet/causal_layer/filter', 'wavenet/dilated_stack/layer0/filter', 'wavenet/dilated_stack/layer0/gate',
            'wavenet/dilated_stack/layer9/filter', 'wavenet/dilated_stack/layer27/dense',
            'wavenet/dilated_stack/layer49/gate', 'wavenet/dilated_stack/layer31/skip',
            'wavenet/dilated_stack/layer20/filter_bias', 'wavenet/postprocessing/postprocess1',
            'wavenet/postprocessing/postprocess2']
    worst_m = worst_e = 0.0
    for k in keys:
        e_m, e_e = l2_rel(got[k], grads_m[k]), l2_rel(got[k], grads_ref[k])
        worst_m, worst_e = max(worst_m, e_m), max(worst_e, e_e)
        assert e_e < GRAD_L2_VS_EXACT, (k, e_e)
        assert e_m < 2e-2, (k, e_m)
    print('cfg2 full size: loss {:.6f} (oracle {:.6f}); worst gradient l2-rel vs exact fp32 oracle {:.2e}, vs '
          'rounding-matched oracle {:.2e}'.format(loss, loss_ref, worst_e, worst_m))


# ----------------------------------------------------------------------------------------- cfg3
def test_cfg3_global_conditioning_377_speakers():
    import wavenet
    kw = dict(DEFAULT_NET, batch_size=2, global_condition_channels=32, global_condition_cardinality=377)
    onet, net = make_pair(O, wavenet, seed=3, **kw)
    audio = _audio(np.random.default_rng(5), 2, 3000)
    gc = [376, 41]
    loss_ref, logits_ref, grads_ref = onet.loss_and_grads(audio, gc)
    loss = float(net.loss(audio, gc))
    got = net.gradients()
    assert abs(loss - loss_ref) <= LOSS_RTOL * abs(loss_ref)
    logits = net.logits(O.mu_law_encode(audio, 256), gc).cpu().numpy()
    assert rel_err(logits, logits_ref) < LOGIT_RTOL
    for k in ['wavenet/embeddings/gc_embedding', 'wavenet/dilated_stack/layer3/gc_filter',
              'wavenet/dilated_stack/layer44/gc_gate', 'wavenet/dilated_stack/layer12/filter']:
        assert l2_rel(got[k], grads_ref[k]) < GRAD_L2_VS_EXACT, k
    emb = got['wavenet/embeddings/gc_embedding']
    used = np.zeros(377, bool)
    used[gc] = True
    assert np.abs(emb[~used]).max() == 0.0 and np.abs(emb[used]).max() > 0.0
    # an id outside the table is refused on the host (the reference's embedding_lookup raises on the CPU)
    with pytest.raises(ValueError):
        net.loss(audio, [377, 0])
    with pytest.raises(ValueError):
        net.loss(audio, [-1, 0])


# ----------------------------------------------------------------------------------------- cfg4
_WRAP_REF = {}


@pytest.mark.parametrize('impl', ['latency', 'throughput'])
def test_cfg4_every_delay_line_wraps(impl):
    """Default params, 6000 teacher-forced steps (> receptive field 5117: the d = 512 rings wrap 11 times)."""
    import wavenet
    from wavenet import _lib
    lib = _lib.load()
    onet, net = make_pair(O, wavenet, seed=4, **DEFAULT_NET)
    ids = np.random.RandomState(4).randint(0, 256, 6000).astype(np.int32)
    checkpoints = [511, 1023, 5116, 5500, 5999]
    ref = _WRAP_REF
    if not ref:      # (both parametrisations share the 6000 oracle steps)
        onet.init_ops()
        for i, s in enumerate(ids):
            p = onet.predict_proba_incremental(int(s))
            if i in checkpoints:
                ref[i] = p
    lib.wn_debug_set_gen_impl(1 if impl == 'latency' else 0)
    try:
        start = 0
        for i in checkpoints:
            p = net.prime(ids[start:i + 1], reset=(start == 0)).cpu().numpy()[0]
            np.testing.assert_allclose(p, ref[i], rtol=5e-3, atol=1e-6, err_msg='step {}'.format(i))
            start = i + 1
    finally:
        lib.wn_debug_set_gen_impl(1)


@pytest.mark.parametrize('kw,streams,steps', [(DEFAULT_NET, 256, 160), (dict(TEST_NET, use_biases=True), 300, 600)],
                         ids=['default_params_256', 'test_net_300_wrapped'])
def test_cfg4_many_streams_vs_oracle(kw, streams, steps):
    """Every stream's distribution after `steps` teacher-forced samples against the batched oracle (the second case
    runs past the receptive field of the small net with 4 streams per CTA and a ragged last CTA)."""
    import wavenet
    onet, net = make_pair(O, wavenet, seed=6, **dict(kw, batch_size=streams))
    ids = np.random.RandomState(6).randint(0, 256, (streams, steps)).astype(np.int32)
    onet.init_ops()
    logits = None
    for i in range(steps):
        logits = onet.predict_proba_incremental(ids[:, i], return_logits=True)
    ref = torch.softmax(torch.tensor(logits, dtype=torch.float64), dim=-1).to(torch.float32).numpy()
    got = net.prime(ids).cpu().numpy()
    assert got.shape == (streams, 256)
    np.testing.assert_allclose(got, ref, rtol=5e-3, atol=1e-6)


@pytest.mark.parametrize('streams', [1, 5])
def test_cfg4_temperature_vs_reference_scaling(streams):
    """generate.py:229-241 at temperature 0.7: the kernel's draws equal np.random.choice on the oracle's scaled
    distribution (teacher-forced with the kernel's own samples)."""
    import wavenet
    kw = dict(TEST_NET, use_biases=True, skip_channels=64)
    onet, net = make_pair(O, wavenet, seed=8, **kw)
    n = 80
    rs = np.random.RandomState(2)
    first = rs.randint(0, 256, streams)
    u = rs.random_sample((streams, n))
    samples = net.generate(n, first, temperature=0.7, uniforms=u).cpu().numpy()
    for s in range(streams):
        onet.init_ops()
        seq = [int(first[s])] + [int(v) for v in samples[s]]
        exact = 0
        for i in range(n):
            p = O.scale_prediction(onet.predict_proba_incremental(seq[i]), 0.7)
            exact += int(O.choice_from_uniform(p, u[s, i]) == samples[s, i])
        assert exact >= n - 1, (s, exact)


def test_generation_refuses_unequal_widths():
    """Both generator kernels assume dilation_channels == residual_channels; other widths must fail loudly."""
    import wavenet
    net = wavenet.WaveNetModel(**dict(TEST_NET, residual_channels=32, dilation_channels=64))
    with pytest.raises(NotImplementedError):
        net.predict_proba_incremental(3)
    with pytest.raises(NotImplementedError):
        net.generate(4, [1])


def test_unaligned_views_are_accepted():
    """An offset view (contiguous, not 16-byte aligned) must work like in the reference."""
    import wavenet
    a = torch.rand(1001, device='cuda') * 2 - 1
    e = wavenet.mu_law_encode(a[1:], 256)
    np.testing.assert_array_equal(e.cpu().numpy(), O.mu_law_encode(a[1:].cpu().numpy(), 256))
    net = wavenet.WaveNetModel(**TEST_NET, seed=1)
    assert float(net.loss(a[1:])) == float(net.loss(a[1:].clone()))


# ----------------------------------------------------------------------------------------- data parallel, 2 GPUs
def test_data_parallel_two_ranks_equals_one_rank_batch_two():
    """DP2 x B1 == 1 GPU x B2 (gradients, 1e-5) and the replicas stay bit-identical after 5 optimizer steps (NCCL)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', str(port),
                        os.path.join(ROOT, 'tests', 'dp_worker.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert 'DP_OK' in r.stdout, r.stdout[-2000:]
