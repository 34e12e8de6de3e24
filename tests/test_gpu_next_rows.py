"""SURVEY section 8f "next" rows on the GPU: generation driver (f1), checkpoint interchange with a real model (f3),
scalar_input front end / summaries (f4)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import wavenet_oracle as O
from wn_helpers import GRAD_L2_VS_EXACT, LOGIT_RTOL, LOSS_RTOL, l2_rel, make_pair, rel_err

GEN_NET = dict(batch_size=1, dilations=[1, 2, 4, 8, 16, 32, 64, 128, 256], filter_width=2, residual_channels=16,
               dilation_channels=16, quantization_channels=128, skip_channels=32)
TEST_NET = dict(batch_size=1, dilations=[1, 2, 4, 8, 16, 32, 64] * 2, filter_width=2, residual_channels=32,
                dilation_channels=32, quantization_channels=256, skip_channels=32)


def _audio(rng, b, t):
    tt = np.arange(t) / 16000.0
    a = 0.3 * np.sin(2 * np.pi * 220 * tt)[None] + 0.3 * np.sin(2 * np.pi * 331 * tt)[None] + 0.1 * rng.standard_normal((b, t))
    return np.clip(a, -1, 1).astype(np.float32)


# ------------------------------------------------------------------------------------------------ f4: scalar_input
@pytest.mark.parametrize('kw,T', [(dict(TEST_NET, scalar_input=True, initial_filter_width=32, use_biases=True, batch_size=2), 700),
                                  (dict(GEN_NET, scalar_input=True, initial_filter_width=5), 400),
                                  (dict(TEST_NET, scalar_input=True, initial_filter_width=32, skip_channels=256,
                                        use_biases=True, global_condition_channels=4, global_condition_cardinality=3,
                                        batch_size=2), 900)], ids=['r32_ifw32', 'r16_ifw5', 'r32_fp16_chain_gc'])
def test_scalar_input_loss_grads_predict_vs_oracle(kw, T):
    """model.py:143-153,570-576,645-648: the causal layer reads the raw float waveform through a width-IFW filter
    (incl. the SAME-padding tap offsets of this snapshot)."""
    import wavenet
    onet, net = make_pair(O, wavenet, seed=5, **kw)
    assert tuple(net.variables['causal_layer']['filter'].shape) == (kw['initial_filter_width'], 1, kw['residual_channels'])
    B = kw['batch_size']
    audio = _audio(np.random.default_rng(2), B, T)
    gc = [2, 0][:B] if kw.get('global_condition_channels') else None
    loss_ref, logits_ref, grads_ref = onet.loss_and_grads(audio, gc)
    loss = float(net.loss(audio, gc))
    assert abs(loss - loss_ref) <= LOSS_RTOL * abs(loss_ref)
    got = net.gradients()
    for k in ('wavenet/causal_layer/filter', 'wavenet/dilated_stack/layer0/filter', 'wavenet/postprocessing/postprocess2'):
        assert l2_rel(got[k], grads_ref[k]) < GRAD_L2_VS_EXACT, k
    if B == 1:
        ids = O.mu_law_encode(audio[0], kw['quantization_channels'])
        np.testing.assert_allclose(net.predict_proba(ids).cpu().numpy(), onet.predict_proba(ids), rtol=5e-3, atol=1e-6)
    with pytest.raises(NotImplementedError):
        net.predict_proba_incremental(3)


def test_summaries_and_histograms():
    import wavenet
    net = wavenet.WaveNetModel(**dict(TEST_NET, use_biases=True, histograms=True), seed=1)
    a = _audio(np.random.default_rng(0), 1, 300)
    s = net.summaries(net.loss(a))
    assert abs(s['loss'] - float(net.loss(a))) < 1e-6 and 'total_loss' not in s
    assert s['layer3_filter'][0].sum() == 2 * 32 * 32 and len(s['layer3_filter'][1]) == 31
    assert 'layer13_biases_skip' in s
    s2 = net.summaries(net.loss(a, l2_regularization_strength=1e-2))
    assert s2['total_loss'] > s2['loss'] and abs(s2['loss'] - s['loss']) < 1e-4


# ------------------------------------------------------------------------------------------------ f3: checkpoints
def test_checkpoint_resume_equals_continuous_training(tmp_path):
    """train.py:104-134 with a real model: save at step 3 (variables + Adam slots), restore into a fresh model, continue --
    the loss curve equals the uninterrupted one; the file holds the reference's variable names."""
    import wavenet
    from wavenet import checkpoint as ck
    kw = dict(TEST_NET, use_biases=True, skip_channels=64)
    audio = _audio(np.random.default_rng(1), 1, 1500)
    net = wavenet.WaveNetModel(**kw, seed=7)
    opt = wavenet.optimizer_factory['adam'](learning_rate=2e-3, momentum=0.9)
    logdir = str(tmp_path / 'train' / 'run')
    curve = []
    for step in range(6):
        loss = net.loss(audio)
        opt.minimize(loss)
        curve.append(float(loss))
        if step == 2:
            ck.save(net, logdir, step, optimizer=opt)
    stored, extra = ck.load_variables(ck.get_checkpoint_state(logdir))
    assert 'wavenet/dilated_stack/layer5/slip_bias' in stored and stored['wavenet/causal_layer/filter'].shape == (2, 256, 32)
    net2 = wavenet.WaveNetModel(**kw, seed=99)
    opt2 = wavenet.optimizer_factory['adam'](learning_rate=2e-3, momentum=0.9)
    assert ck.load(net2, logdir, optimizer=opt2) == 2
    resumed = []
    for step in range(3, 6):
        loss = net2.loss(audio)
        opt2.minimize(loss)
        resumed.append(float(loss))
    np.testing.assert_allclose(resumed, curve[3:], rtol=2e-4)
    # the TF auto-named form of the same checkpoint restores too (generate.py:176-182 on this snapshot's files)
    path = str(tmp_path / 'model.ckpt-77')
    ck.save_variables(path, ck.to_tf_autonames(net.state_dict()))
    net3 = wavenet.WaveNetModel(**kw, seed=5)
    ck.restore(net3, path)
    assert torch.equal(net3.flat_params, net.flat_params)
    assert ck.step_of(path) == 77


# ------------------------------------------------------------------------------------------------ f1: generation driver
def test_generation_driver_fast_equals_slow_and_reference_loop(tmp_path):
    """generate.py:187-272 as a library call.  With the global numpy stream seeded the same way, fast generation (the
    persistent kernel drawing from uploaded uniforms), slow generation (predict_proba + np.random.choice per sample) and
    the reference loop restated on the oracle produce the same waveform."""
    import wavenet
    from wavenet import generation
    onet, net = make_pair(O, wavenet, seed=3, **dict(GEN_NET, use_biases=True))
    n = 40
    np.random.seed(11)
    fast, audio_fast = generation.generate(net, samples=n, temperature=0.8, fast_generation=True)
    np.random.seed(11)
    slow, _ = generation.generate(net, samples=n, temperature=0.8, fast_generation=False, window=8000)
    assert len(fast) == n + 1 and fast[0] == slow[0]
    assert sum(int(a != b) for a, b in zip(fast, slow)) == 0
    # the reference's loop on the oracle
    np.random.seed(11)
    wave = np.random.randint(128, size=(1,)).tolist()
    onet.init_ops()
    for _ in range(n):
        p = O.scale_prediction(onet.predict_proba_incremental(wave[-1]), 0.8)
        wave.append(int(np.random.choice(np.arange(128), p=p)))
    assert sum(int(a != b) for a, b in zip(fast, wave)) <= 1
    assert audio_fast.shape == (n + 1,) and np.abs(audio_fast).max() <= 1.0
    np.testing.assert_array_equal(audio_fast, O.mu_law_decode(np.asarray(fast), 128))


def test_generation_driver_seed_priming_and_save_every(tmp_path):
    import wavenet
    from wavenet import audio_reader, generation
    onet, net = make_pair(O, wavenet, seed=4, **dict(GEN_NET, use_biases=True))
    # a wav seed: 0.1 s of silence, a tone, silence
    t = np.arange(3000) / 16000.0
    x = np.concatenate([np.zeros(2600), 0.4 * np.sin(2 * np.pi * 300 * t), np.zeros(2600)]).astype(np.float32)
    seed_path, out_path = str(tmp_path / 'seed.wav'), str(tmp_path / 'out.wav')
    generation.write_wav(x, 16000, seed_path)
    seed = generation.create_seed(seed_path, 16000, 128, window_size=2000)
    trimmed = audio_reader.trim_silence(x, generation.SILENCE_THRESHOLD)
    assert len(seed) == 2000 and seed == O.mu_law_encode(trimmed, 128)[:2000].tolist()
    # window = 200: the first len(seed) - 201 samples are primed (generate.py:204), then the loop continues from the last one
    np.random.seed(5)
    wave, audio = generation.generate(net, samples=24, wav_seed=seed_path, window=200, save_every=8, wav_out_path=out_path,
                                      sample_rate=16000)
    full_seed = generation.create_seed(seed_path, 16000, 128)
    assert wave[:len(full_seed)] == full_seed and len(wave) == len(full_seed) + 24
    np.random.seed(5)
    onet.init_ops()
    for s in full_seed[:-201]:
        onet.predict_proba_incremental(int(s))
    ref = list(full_seed)
    for _ in range(24):
        p = O.scale_prediction(onet.predict_proba_incremental(ref[-1]), 1.0)
        ref.append(int(np.random.choice(np.arange(128), p=p)))
    assert sum(int(a != b) for a, b in zip(wave, ref)) <= 1
    written = audio_reader.load_wav(out_path, 16000)
    np.testing.assert_array_equal(written, audio)
