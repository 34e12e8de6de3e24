"""Parity of the reference-facing Python API (WaveNetModel / ops) on sm_100a against the CPU
oracle, plus ports of the reference's own model / generation tests."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import wavenet_oracle as O
from wn_helpers import (GRAD_L2_VS_EXACT, GRAD_RTOL, LOGIT_RTOL, LOGIT_RTOL_16BIT, LOSS_RTOL, l2_rel, make_pair, matched_oracle,
                        rel_err)

TEST_NET = dict(batch_size=1, dilations=[1, 2, 4, 8, 16, 32, 64] * 2, filter_width=2, residual_channels=32,
                dilation_channels=32, quantization_channels=256, skip_channels=32)          # test_model.py:190-199
GEN_NET = dict(batch_size=1, dilations=[1, 2, 4, 8, 16, 32, 64, 128, 256], filter_width=2, residual_channels=16,
               dilation_channels=16, quantization_channels=128, skip_channels=32)           # test_generation.py:10-17
DEFAULT_NET = dict(batch_size=1, dilations=[2 ** i for i in range(10)] * 5, filter_width=2, residual_channels=32,
                   dilation_channels=32, quantization_channels=256, skip_channels=512, use_biases=True)


def make_sine_waves(global_conditioning):
    """test/test_model.py:29-58 (random.randint leading silence fixed by seeding)."""
    import random
    times = np.arange(0.0, 0.5, 1.0 / 2000.0)
    if global_conditioning:
        random.seed(42)
        lead = random.randint(10, 128)
        amp = np.zeros((3, len(times)))
        tt = times[lead:] - lead / 2000.0
        amp[0, lead:] = 0.6 * np.sin(tt * 2.0 * np.pi * 155.56)
        amp[1, lead:] = 0.5 * np.sin(tt * 2.0 * np.pi * 196.00)
        amp[2, lead:] = 0.4 * np.sin(tt * 2.0 * np.pi * 233.08)
        return amp, np.array([[0], [1], [2]])
    amp = (np.sin(times * 2.0 * np.pi * 155.56) / 3.0 + np.sin(times * 2.0 * np.pi * 196.00) / 3.0 +
           np.sin(times * 2.0 * np.pi * 233.08) / 3.0)
    return amp, None


def _audio(rng, b, t):
    tt = np.arange(t) / 16000.0
    a = 0.3 * np.sin(2 * np.pi * 220 * tt)[None] + 0.3 * np.sin(2 * np.pi * 331 * tt)[None] + \
        0.1 * rng.standard_normal((b, t))
    return np.clip(a, -1, 1).astype(np.float32)


CASES = {
    'test_net': (dict(TEST_NET), 1000, None),
    'test_net_biases': (dict(TEST_NET, use_biases=True), 1000, None),
    'gc_batch3': (dict(TEST_NET, batch_size=3, use_biases=True, skip_channels=256, global_condition_channels=3,
                       global_condition_cardinality=3), 700, [0, 1, 2]),
    'gc_xavier_table': (dict(TEST_NET, batch_size=2, use_biases=True, global_condition_channels=8,
                             global_condition_cardinality=5), 300, [4, 1]),
    'residual_postproc': (dict(TEST_NET, batch_size=2, use_biases=True, skip_channels=64, residual_postproc=True),
                          500, None),
    'gen_net_r16': (dict(GEN_NET, batch_size=2, use_biases=True), 777, None),
    'default_params_short': (dict(DEFAULT_NET), 6000, None),
    # R, D multiples of 64: wide blocks in 16-bit storage (block_wide16.cu); other widths: GEMM-built fp32 blocks (block_generic.cu)
    'wide_r64_d64': (dict(TEST_NET, batch_size=2, residual_channels=64, dilation_channels=64, skip_channels=64,
                          use_biases=True), 600, None),
    'r32_d64_gc': (dict(TEST_NET, batch_size=2, residual_channels=32, dilation_channels=64, use_biases=True,
                        global_condition_channels=8, global_condition_cardinality=5), 500, [3, 0]),
    'scaled_r128_short': (dict(batch_size=1, dilations=[2 ** i for i in range(10)] * 2, filter_width=2,
                               residual_channels=128, dilation_channels=128, quantization_channels=256,
                               skip_channels=512, use_biases=True), 3000, None),
    # BASELINE config 5 (scaled net: res / dil 128, skip 512, 4 x dilations 1..512) on short windows: two batch elements
    # (the dilated operand must read zeros, not the neighbouring element), T past the receptive field 4093, and T < d
    'cfg5_scaled_net_b2': (dict(batch_size=2, dilations=[2 ** i for i in range(10)] * 4, filter_width=2,
                                residual_channels=128, dilation_channels=128, quantization_channels=256,
                                skip_channels=512, use_biases=True), 4500, None),
    'cfg5_scaled_net_T300_gc': (dict(batch_size=2, dilations=[2 ** i for i in range(10)] * 4, filter_width=2,
                                     residual_channels=128, dilation_channels=128, quantization_channels=256,
                                     skip_channels=512, use_biases=True, global_condition_channels=16,
                                     global_condition_cardinality=7), 300, [5, 2]),
    'narrow_r8_d12': (dict(TEST_NET, residual_channels=8, dilation_channels=12, skip_channels=20), 300, None),
    # BASELINE config 3: default params + global conditioning on 377 speakers (32 channels)
    'cfg3_gc377': (dict(DEFAULT_NET, batch_size=2, global_condition_channels=32, global_condition_cardinality=377),
                   2500, [376, 7]),
}
# The "matched" oracle rounds operands to TF32 where the first-generation kernels did; the production kernels round to
# fp16 / split fp16 at other positions, so it is no tighter than the exact oracle (both 1-4e-2: rounding noise amplified
# by back-propagation at random initialisation, tests/test_oracle_network.py).  It is reported, not asserted; kernel
# LOGIC is pinned (a) per kernel on the production kernels at 4e-3 max-norm (test_gpu_kernels.py::test_block_fwd_bwd),
# (b) by the launch-structure equivalence test at the bottom of this file (1e-4 max-norm, also at the benchmarked size).


@pytest.mark.parametrize('case', sorted(CASES))
def test_loss_logits_grads_vs_oracle(case):
    import wavenet
    kw, T, gc = CASES[case]
    onet, net = make_pair(O, wavenet, seed=1, **kw)
    rng = np.random.default_rng(7)
    audio = _audio(rng, kw['batch_size'], T)
    loss_ref, logits_ref, grads_ref = onet.loss_and_grads(audio, gc)
    loss = net.loss(audio, gc)
    ids = O.mu_law_encode(audio, kw['quantization_channels'])
    logits = net.logits(ids, gc).cpu().numpy()
    assert abs(float(loss) - loss_ref) <= LOSS_RTOL * abs(loss_ref), (float(loss), loss_ref)
    # 16-bit activation storage (R, D multiples of 64, BASELINE config 5 "bf16 training"): the residual stream is rounded to
    # fp16 once per layer, 2^-11 relative each time; measured 1.2-1.3e-3 of the logit scale after 14-40 layers
    wide16 = kw['residual_channels'] % 64 == 0 and kw['dilation_channels'] % 64 == 0
    assert rel_err(logits, logits_ref) < (LOGIT_RTOL_16BIT if wide16 else LOGIT_RTOL)
    got = net.gradients()
    # gradients: tight against the arithmetic-matched oracle, norm-wise against the exact one
    _, _, grads_m = matched_oracle(O, onet, **kw).loss_and_grads(audio, gc)
    worst, worst_l2, bad = 0.0, 0.0, []
    n_last = len(kw['dilations']) - 1
    for k, g in grads_ref.items():
        if k == 'wavenet/dilated_stack/layer{}/dense'.format(n_last) or \
                k == 'wavenet/dilated_stack/layer{}/dense_bias'.format(n_last):
            assert np.abs(got[k]).max() == 0.0       # no gradient reaches them (App. A5)
            continue
        e = l2_rel(got[k], grads_m[k])
        e2 = l2_rel(got[k], g)
        cos = float(np.dot(got[k].ravel().astype(np.float64), g.ravel().astype(np.float64)) /
                    max(np.linalg.norm(got[k]) * np.linalg.norm(g), 1e-300))
        worst, worst_l2 = max(worst, e), max(worst_l2, e2)
        if e2 >= GRAD_L2_VS_EXACT or cos < 0.998:
            bad.append((k, e2, cos, e))
    print('case {} loss {:.6f} ref {:.6f} logits rel {:.2e} | grads: worst l2-rel vs exact {:.2e} (vs TF32-emulating '
          'oracle {:.2e})'.format(case, float(loss), loss_ref, rel_err(logits, logits_ref), worst_l2, worst))
    assert not bad, bad[:8]


def test_loss_input_shapes_and_l2():
    import wavenet
    onet, net = make_pair(O, wavenet, seed=2, **dict(TEST_NET, use_biases=True))
    a = _audio(np.random.default_rng(0), 1, 400)
    l1 = float(net.loss(a[0]))                   # [T]
    l2 = float(net.loss(a))                      # [B,T]
    l3 = float(net.loss(a[:, :, None]))          # [B,T,1]
    assert l1 == l2 == l3
    ref = float(onet.loss(a, None, 1e-3).detach())
    got = float(net.loss(a, l2_regularization_strength=1e-3))
    assert abs(got - ref) <= LOSS_RTOL * abs(ref)


def test_predict_proba_vs_oracle():
    import wavenet
    onet, net = make_pair(O, wavenet, seed=3, **dict(GEN_NET, use_biases=True))
    np.random.seed(0)
    data = np.random.randint(128, size=1000)      # test_generation.py:19-32
    proba = net.predict_proba(data).cpu().numpy()
    assert proba.shape == (128,)
    assert np.all((proba >= 0) & (proba <= 127))
    assert abs(proba.sum() - 1) < 1e-5
    np.testing.assert_allclose(proba, onet.predict_proba(data), rtol=5e-3, atol=1e-6)


def test_generate_fast_and_compare_simple_fast():
    """test_generation.py:34-72, with the priming loop made non-trivial (600 > receptive field)."""
    import wavenet
    for biases in (False, True):
        onet, net = make_pair(O, wavenet, seed=4, **dict(GEN_NET, use_biases=biases))
        np.random.seed(0)
        data = np.random.randint(128, size=600)
        for op in net.init_ops:
            op()
        p0 = net.predict_proba_incremental(int(data[0])).cpu().numpy()
        assert p0.shape == (128,) and np.all((p0 >= 0) & (p0 <= 127))
        p0b = net.predict_proba_incremental(int(data[0])).cpu().numpy()     # no push in between: same state
        np.testing.assert_array_equal(p0, p0b)
        for x in data[:-1]:
            net.predict_proba_incremental(int(x))
            for op in net.push_ops:
                op()
        proba_fast = net.predict_proba_incremental(int(data[-1])).cpu().numpy()
        proba = net.predict_proba(data).cpu().numpy()
        np.testing.assert_allclose(proba, proba_fast, rtol=5e-3, atol=1e-6)
        # and both against the oracle's incremental generator
        onet.init_ops()
        for x in data[:-1]:
            onet.predict_proba_incremental(x)
        np.testing.assert_allclose(proba_fast, onet.predict_proba_incremental(data[-1]), rtol=5e-3, atol=1e-6)
        # one-launch priming gives the same state as the per-sample protocol
        proba_primed = net.prime(data).cpu().numpy()[0]
        np.testing.assert_allclose(proba_primed, proba_fast, rtol=1e-5, atol=1e-8)


def test_incremental_refuses_unsupported():
    import wavenet
    net = wavenet.WaveNetModel(**dict(TEST_NET, filter_width=3))
    with pytest.raises(NotImplementedError):
        net.predict_proba_incremental(1)                       # model.py:597-599
    net = wavenet.WaveNetModel(**dict(TEST_NET, scalar_input=True, initial_filter_width=4))
    with pytest.raises(NotImplementedError):
        net.predict_proba_incremental(1)                       # model.py:601-603


def test_generate_batched_gc_streams_vs_oracle():
    """256-stream style batch (here 5 streams, gc per stream): every step's distribution, teacher-forced
    with the kernel's own draws, matches the oracle; draws equal np.random.choice on those distributions."""
    import wavenet
    kw = dict(TEST_NET, use_biases=True, skip_channels=64, global_condition_channels=4, global_condition_cardinality=6)
    onet, net = make_pair(O, wavenet, seed=5, **kw)
    streams, n = 5, 40
    rng = np.random.RandomState(1)
    first = rng.randint(0, 256, streams)
    gc = rng.randint(0, 6, streams)
    u = rng.random_sample((streams, n))
    samples = net.generate(n, first, global_condition=gc, uniforms=u).cpu().numpy()
    assert samples.shape == (streams, n) and samples.min() >= 0 and samples.max() < 256
    for s in range(streams):
        onet.init_ops()
        seq = [int(first[s])] + [int(v) for v in samples[s]]
        exact = 0
        for i in range(n):
            p_ref = onet.predict_proba_incremental(seq[i], int(gc[s]))
            if O.choice_from_uniform(p_ref, u[s, i]) == samples[s, i]:
                exact += 1
        assert exact >= n - 1      # a draw may differ only when u sits within float noise of a cdf edge
    # the same draws come out of the step-by-step API + wn_sample (identical kernels => identical bits)
    from wavenet import _lib
    import ctypes as C
    lib = _lib.load()
    for op in net.init_ops:
        op(streams)
    cur = torch.as_tensor(first.astype(np.int32), device='cuda')
    for i in range(5):
        proba = torch.empty((streams, 256), device='cuda')
        g = net._gen
        rc = lib.wn_gen_run(C.byref(net._cfg), _lib.ptr(net.flat_params), _lib.ptr(g['state']), streams,
                            _lib.ptr(cur), None, _lib.ptr(torch.as_tensor(gc.astype(np.int32), device='cuda')), None,
                            1, 1.0, 1, None, _lib.ptr(proba), _lib.stream_ptr())
        assert rc == 0
        out = torch.empty(streams, dtype=torch.int32, device='cuda')
        uu = torch.as_tensor(u[:, i].copy(), device='cuda')
        assert lib.wn_sample(_lib.ptr(proba), _lib.ptr(uu), streams, 256, _lib.ptr(out), _lib.stream_ptr()) == 0
        np.testing.assert_array_equal(out.cpu().numpy(), samples[:, i])
        cur = out


@pytest.mark.parametrize('kw,gc', [
    (dict(TEST_NET, use_biases=True, skip_channels=64, global_condition_channels=4, global_condition_cardinality=6), 3),
    (dict(DEFAULT_NET), None)], ids=['test_net_gc', 'default_params'])
def test_generate_latency_kernel(kw, gc):
    """One stream = the latency-mode kernel (layer-per-warp chain CTAs + resident post-processing CTAs + sampler
    CTA talking through tagged words).  Its distributions, teacher-forced with its own draws, match the oracle;
    its draws are np.random.choice on those distributions; the delay lines / header it leaves behind are the
    ones the throughput-mode kernel continues from."""
    import wavenet
    from wavenet import _lib
    lib = _lib.load()
    onet, net = make_pair(O, wavenet, seed=7, **kw)
    n = 48 if len(kw['dilations']) > 20 else 150
    rs = np.random.RandomState(3)
    u = rs.random_sample((1, n))
    first = [int(rs.randint(0, 256))]
    lib.wn_debug_set_gen_impl(1)
    samples, proba = net.generate(n, first, global_condition=gc, uniforms=u, return_proba=True)
    samples = samples.cpu().numpy()[0]
    assert samples.min() >= 0 and samples.max() < 256
    onet.init_ops()
    seq = first + [int(v) for v in samples]
    exact = 0
    p_ref = None
    for i in range(n):
        p_ref = onet.predict_proba_incremental(seq[i], gc) if gc is not None else onet.predict_proba_incremental(seq[i])
        exact += int(O.choice_from_uniform(p_ref, u[0, i]) == samples[i])
    assert exact >= n - 1          # a draw may differ only when u sits within float noise of a cdf edge
    np.testing.assert_allclose(proba.cpu().numpy()[0], p_ref, rtol=5e-3, atol=1e-6)
    # continue a few more steps with the throughput-mode kernel from the state the latency kernel left ...
    u2 = rs.random_sample((1, 8))
    lib.wn_debug_set_gen_impl(0)
    try:
        cont_v1 = net.generate(8, [int(samples[-1])], global_condition=gc, uniforms=u2, reset=False).cpu().numpy()[0]
        # ... and the same with the latency kernel after replaying the identical history through the v1 kernel
        net.prime(np.concatenate([first, samples[:-1]]).astype(np.int32), global_condition=gc)
    finally:
        lib.wn_debug_set_gen_impl(1)
    cont_lat = net.generate(8, [int(samples[-1])], global_condition=gc, uniforms=u2, reset=False).cpu().numpy()[0]
    assert (cont_v1 != cont_lat).sum() <= 1
    # priming (forced ids, no draws) through both kernels
    ids = np.concatenate([first, samples[:-1]]).astype(np.int32)
    p_lat = net.prime(ids, global_condition=gc).cpu().numpy()[0]
    lib.wn_debug_set_gen_impl(0)
    try:
        p_v1 = net.prime(ids, global_condition=gc).cpu().numpy()[0]
    finally:
        lib.wn_debug_set_gen_impl(1)
    np.testing.assert_allclose(p_lat, p_v1, rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(p_lat, p_ref, rtol=5e-3, atol=1e-6)


def test_generate_temperature_and_many_streams():
    import wavenet
    net = wavenet.WaveNetModel(**dict(TEST_NET, use_biases=True), seed=3)
    streams = 300                                    # > 2 * SM count -> 4 streams per CTA, ragged tail
    first = np.random.RandomState(0).randint(0, 256, streams)
    a = net.generate(16, first, temperature=0.7, seed=5).cpu().numpy()
    b = net.generate(16, first, temperature=0.7, seed=5).cpu().numpy()
    np.testing.assert_array_equal(a, b)             # deterministic given the uniforms
    assert a.shape == (streams, 16) and a.min() >= 0 and a.max() < 256
    # streams are independent: a subset generated alone gives the same samples
    u = np.stack([np.random.RandomState(5 + s).random_sample(16) for s in range(streams)])
    c = net.generate(16, first[:3], temperature=0.7, uniforms=u[:3]).cpu().numpy()
    np.testing.assert_array_equal(a[:3], c)


@pytest.mark.parametrize('opt_name,lr,biases,skip', [('sgd', 0.02, False, 32), ('sgd', 0.02, True, 32),
                                                      ('rmsprop', 0.001, False, 256), ('adam', 0.002, True, 32)])
def test_end_to_end_training_sine(opt_name, lr, biases, skip):
    """test/test_model.py:222-282: loss after 400 its < 0.1 and < 2% of the initial loss."""
    import wavenet
    audio, _ = make_sine_waves(False)
    net = wavenet.WaveNetModel(**dict(TEST_NET, use_biases=biases, skip_channels=skip), seed=42)
    opt = wavenet.optimizer_factory[opt_name](learning_rate=lr, momentum=0.95)
    initial = float(net.loss(audio))
    loss = None
    for _ in range(400):
        loss = net.loss(audio)
        opt.minimize(loss)
    final = float(net.loss(audio))
    assert initial > 0.1 and final < 0.1 and final / initial < 0.02, (initial, final)
    # generation check (test_model.py:137-172): tone power dominates the spectrum of generated audio
    if opt_name == 'rmsprop':
        samples = net.generate(1000, [128], seed=0).cpu().numpy()[0]
        wave = wavenet.mu_law_decode(samples[255:], 256).cpu().numpy()
        power = np.abs(np.fft.fft(wave)) ** 2
        freqs = np.fft.fftfreq(wave.size, 1.0 / 2000.0)
        sel = (freqs >= 0) & (freqs <= 500)
        power, freqs = power[sel], freqs[sel]
        tone = sum(power[np.abs(freqs - f).argmin()] for f in (155.56, 196.00, 233.08))
        assert tone > 0.7 * power.sum()


def test_end_to_end_training_gc():
    """test/test_model.py:384-405: 3 speakers, one tone each, global conditioning."""
    import wavenet
    audio, ids = make_sine_waves(True)
    net = wavenet.WaveNetModel(**dict(TEST_NET, batch_size=3, use_biases=True, skip_channels=256,
                                      global_condition_channels=3, global_condition_cardinality=3), seed=42)
    opt = wavenet.optimizer_factory['sgd'](learning_rate=0.01, momentum=0.95)
    initial = float(net.loss(audio, ids))
    for _ in range(1000):
        opt.minimize(net.loss(audio, ids))
    final = float(net.loss(audio, ids))
    assert initial > 0.1 and final < 0.1 and final / initial < 0.02, (initial, final)


def test_train_step_graph_matches_eager():
    import wavenet
    kw = dict(TEST_NET, use_biases=True)
    a = _audio(np.random.default_rng(0), 1, 2000)
    nets = [wavenet.WaveNetModel(**kw, seed=9) for _ in range(2)]
    opts = [wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9) for _ in range(2)]
    step = wavenet.TrainStep(nets[0], opts[0], 1, 2000)
    for _ in range(3):
        l_graph = float(step(a))
        l_eager = nets[1].loss(a)
        opts[1].minimize(l_eager)
        assert abs(l_graph - float(l_eager)) < 1e-4 * abs(l_graph)
    # (the two runs order their fp32 atomics differently, and Adam normalises tiny gradients)
    np.testing.assert_allclose(nets[0].flat_params.cpu().numpy(), nets[1].flat_params.cpu().numpy(), atol=2e-4)


@pytest.mark.parametrize('case', ['test_net_biases', 'gc_batch3', 'residual_postproc', 'default_params_short'])
def test_tcgen05_path_matches_mma_sync_path(case):
    """The tcgen05 kernels (TMA + UMMA + TMEM GEMMs and residual blocks) against the independent mma.sync
    implementation of the same arithmetic: loss, logits and every gradient."""
    import wavenet
    from wavenet import _lib
    lib = _lib.load()
    kw, T, gc = CASES[case]
    audio = _audio(np.random.default_rng(3), kw['batch_size'], T)
    ids = O.mu_law_encode(audio, kw['quantization_channels'])
    out = {}
    try:
        for name, flag in (('umma', 0), ('mma', 1)):
            assert lib.wn_debug_set_impl(flag, flag) == 0
            net = wavenet.WaveNetModel(**kw, seed=11)
            out[name] = (float(net.loss(audio, gc)), net.logits(ids, gc).cpu().numpy(), net.gradients())
    finally:
        lib.wn_debug_set_impl(0, 0)
    assert abs(out['umma'][0] - out['mma'][0]) < 2e-4 * abs(out['mma'][0])
    assert rel_err(out['umma'][1], out['mma'][1]) < LOGIT_RTOL
    worst = 0.0
    for k, g in out['mma'][2].items():
        e = l2_rel(out['umma'][2][k], g) if np.abs(g).max() > 0 else float(np.abs(out['umma'][2][k]).max())
        worst = max(worst, e)
        assert e < 2e-2, (k, e)
    print('case {}: tcgen05 vs mma.sync worst gradient l2-rel {:.2e}'.format(case, worst))


def test_default_params_full_size_properties():
    """BASELINE config (default wavenet_params.json, T = 100000): size-independent checks."""
    import wavenet
    net = wavenet.WaveNetModel(**DEFAULT_NET, seed=0)
    rng = np.random.default_rng(0)
    a = _audio(rng, 1, 100000)
    loss = float(net.loss(a))
    assert math.isfinite(loss) and abs(loss - math.log(256)) < 0.5       # random init ~ ln Q
    g1 = net.flat_grads.clone()
    assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
    # batching invariance: the same window twice in a batch gives the same mean loss and gradients
    net2 = wavenet.WaveNetModel(**dict(DEFAULT_NET, batch_size=2), seed=0)
    net2.load_state_dict(net.state_dict())
    loss2 = float(net2.loss(np.concatenate([a, a])))
    assert abs(loss2 - loss) < 1e-4 * abs(loss)
    assert rel_err(net2.flat_grads.cpu().numpy(), g1.cpu().numpy()) < 1e-3
    # causality: perturbing the last 1000 samples leaves the logits of the first 99000 unchanged
    ids = O.mu_law_encode(a, 256)
    ids2 = ids.copy()
    ids2[:, -1000:] = (ids2[:, -1000:] + 17) % 256
    l_a = net.logits(ids)[0, :99000]
    l_b = net.logits(ids2)[0, :99000]
    assert torch.equal(l_a, l_b)


_VARIANT_SCRIPT = r"""
import json, os, sys
sys.path.insert(0, os.path.join({root!r}, 'tensorflow-wavenet_b200'))
import numpy as np, wavenet
p = json.load(open(os.path.join({root!r}, 'tensorflow-wavenet_b200', 'wavenet_params.json')))
B, T = int(sys.argv[2]), int(sys.argv[3])
net = wavenet.WaveNetModel(batch_size=B, dilations=p['dilations'], filter_width=2, residual_channels=32, dilation_channels=32,
                           quantization_channels=256, skip_channels=512, use_biases=True, global_condition_channels=16,
                           global_condition_cardinality=5, seed=3)
rng = np.random.default_rng(5)
a = np.clip(0.5 * np.sin(np.arange(T) * 0.03)[None] + 0.2 * rng.standard_normal((B, T)), -1, 1).astype(np.float32)
loss = float(net.loss(a, [1, 4][:B]))
np.savez(sys.argv[1], loss=loss, grads=net.flat_grads.cpu().numpy())
"""


def _run_variants(tmp_path, variants, B, T):
    import subprocess
    import sys
    script = tmp_path / 'variant.py'
    script.write_text(_VARIANT_SCRIPT.format(root=ROOT))
    out = {}
    for name, env in variants.items():
        path = str(tmp_path / (name + '.npz'))
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, str(script), path, str(B), str(T)], env=e, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[name] = np.load(path)
    return out


@pytest.mark.parametrize('B,T', [(2, 20011), (1, 100000)], ids=['b2_t20011', 'benchmark_shape'])
def test_persistent_kernels_match_per_layer_launches(tmp_path, B, T):
    """Launch structure and arithmetic variants of the training step on default params (the second case is the benchmarked
    shape: 782 tiles per layer, every CTA of the persistent kernels busy).  The switches are read once per process, hence
    the subprocesses.
      * default twice: the persistent kernels claim their tiles dynamically, so two runs order their atomics differently
        and nothing else -- a race (a tile read before it was published) would show up as a difference;
      * first-generation backward (per-layer TF32 kernels) with the forward layers and weight gradients as per-layer
        launches vs as persistent kernels: same arithmetic, different launch structure;
      * fp16 backward chain (default) vs the first-generation TF32 backward, fp16 vs TF32 post-processing GEMMs: the same
        11-bit operand precision with different roundings."""
    variants = {'default': {}, 'default_again': {},
                'bwd_tf32': {'WN_BWD_CHAIN': '0'},
                'bwd_tf32_per_layer': {'WN_BWD_CHAIN': '0', 'WN_FWD_CHAIN': '0', 'WN_WGRAD_PER_LAYER': '1'},
                'tf32_gemms': {'WN_FWD_GEMM': 'tf32'}}
    out = _run_variants(tmp_path, variants, B, T)
    ref = out['default']
    assert math.isfinite(float(ref['loss'])) and np.isfinite(ref['grads']).all()
    assert float(out['default_again']['loss']) == float(ref['loss'])
    assert rel_err(out['default_again']['grads'], ref['grads']) < 2e-5
    # same arithmetic, different launch structure: only the order of the split-K / atomic accumulations differs
    old = out['bwd_tf32']
    assert abs(float(out['bwd_tf32_per_layer']['loss']) - float(old['loss'])) <= 1e-6 * abs(float(old['loss']))
    assert rel_err(out['bwd_tf32_per_layer']['grads'], old['grads']) < 1e-4
    # fp16 split rows vs TF32 in the block backward, fp16 vs TF32 operand copies in the GEMMs: 11-bit mantissas, different roundings
    assert float(old['loss']) == float(ref['loss'])
    assert rel_err(old['grads'], ref['grads']) < 2e-2
    assert abs(float(out['tf32_gemms']['loss']) - float(ref['loss'])) <= LOSS_RTOL * abs(float(ref['loss']))
    assert rel_err(out['tf32_gemms']['grads'], ref['grads']) < 2e-2


def test_wide_blocks_16bit_storage_match_fp32_gemm_built_blocks():
    """R = D = 128: the 16-bit-storage blocks (block_wide16.cu, default) against the fp32 / TF32 GEMM-built blocks
    (block_generic.cu, WN_WIDE16=0, read once per process -> a child process): loss 1e-3, flat gradient 5e-2 L2."""
    import subprocess
    import sys
    import tempfile
    code = r'''
import os, sys
sys.path.insert(0, os.path.join({root!r}, 'tensorflow-wavenet_b200'))
import numpy as np, wavenet
net = wavenet.WaveNetModel(batch_size=2, dilations=[1, 2, 4, 8, 16, 32, 64, 128] * 2, filter_width=2, residual_channels=128,
                           dilation_channels=128, quantization_channels=256, skip_channels=128, use_biases=True, seed=5)
a = np.clip(0.4 * np.sin(np.arange(2 * 1100) * 0.03).reshape(2, 1100) + 0.05 * np.random.default_rng(1).standard_normal((2, 1100)), -1, 1)
loss = float(net.loss(a.astype(np.float32)))
np.save(sys.argv[1], np.concatenate([[loss], net.flat_grads.cpu().numpy().ravel()]))
'''.format(root=ROOT)
    out = []
    with tempfile.TemporaryDirectory() as td:
        for i, env in enumerate(({}, {'WN_WIDE16': '0'})):
            path = os.path.join(td, 'r%d.npy' % i)
            subprocess.run([sys.executable, '-c', code, path], check=True, env=dict(os.environ, **env), timeout=300)
            out.append(np.load(path))
    assert abs(out[0][0] - out[1][0]) <= 1e-3 * abs(out[1][0]), (out[0][0], out[1][0])
    assert l2_rel(out[0][1:], out[1][1:]) < 5e-2
    assert not np.array_equal(out[0][1:], out[1][1:])      # (two different implementations did run)
