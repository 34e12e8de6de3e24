"""Self-consistency pins of the network oracle that the reference's own tests hold
(test/test_generation.py:50-72 naive == incremental; test/test_model.py:275-282 loss thresholds)
plus the documented TF semantics."""
import math

import numpy as np
import pytest
import torch

import wavenet_oracle as O


def test_naive_equals_incremental_beyond_receptive_field():
    net = O.OracleWaveNet(1, [1, 2, 4, 8, 16, 32, 64, 128, 256], 2, 16, 16, 32, quantization_channels=128,
                          use_biases=True, bias_scale=0.1, dtype=torch.float64)
    np.random.seed(0)
    data = np.random.randint(128, size=600)
    net.init_ops()
    for s in data[:-1]:
        net.predict_proba_incremental(s)
    np.testing.assert_allclose(net.predict_proba_incremental(data[-1]), net.predict_proba(data), atol=1e-7)


def test_forward_without_push_leaves_state():
    net = O.OracleWaveNet(1, [1, 2, 4], 2, 16, 16, 32, quantization_channels=64)
    net.init_ops()
    net.predict_proba_incremental(3)
    a = net.predict_proba_incremental(5, push=False)
    b = net.predict_proba_incremental(5, push=False)
    np.testing.assert_array_equal(a, b)


def test_initial_loss_is_log_q():
    net = O.OracleWaveNet(1, [1, 2, 4, 8] * 2, 2, 32, 32, 32)
    audio = np.random.default_rng(1).uniform(-1, 1, 400).astype(np.float32)
    loss = float(net.loss(audio))
    assert abs(loss - math.log(256) * 399 / 400) < 0.2


def test_tf_xent_gradient_quirk_last_row():
    """The all-zero label row contributes no loss but TF backprop = softmax - labels (App. A4)."""
    logits = torch.randn(3, 5, dtype=torch.float64, requires_grad=True)
    labels = torch.zeros(3, 5, dtype=torch.float64)
    labels[0, 1] = labels[1, 4] = 1.0
    per = O._TFSoftmaxXent.apply(logits, labels)
    assert per[2].item() == 0.0
    per.sum().backward()
    np.testing.assert_allclose(logits.grad[2].numpy(), torch.softmax(logits[2], -1).detach().numpy(), atol=1e-12)


def test_faithful_and_closed_form_paths_agree():
    net = O.OracleWaveNet(2, [1, 2, 4, 1, 2, 4], 2, 32, 32, 64, use_biases=True, bias_scale=0.1,
                          global_condition_channels=4, global_condition_cardinality=3, residual_postproc=True,
                          dtype=torch.float64)
    a = np.random.default_rng(2).uniform(-1, 1, (2, 50)).astype(np.float32)
    l1, lg1, g1 = net.loss_and_grads(a, [0, 2])
    net.faithful = False
    l2, lg2, g2 = net.loss_and_grads(a, [0, 2])
    assert abs(l1 - l2) < 1e-10
    for k in g1:
        np.testing.assert_allclose(g1[k], g2[k], atol=1e-6)


def make_sine_waves():
    # test/test_model.py:29-58 (non-gc branch)
    times = np.arange(0.0, 0.5, 1.0 / 2000.0)
    return (np.sin(times * 2.0 * np.pi * 155.56) / 3.0 + np.sin(times * 2.0 * np.pi * 196.00) / 3.0 +
            np.sin(times * 2.0 * np.pi * 233.08) / 3.0)


@pytest.mark.timeout(600)
def test_sine_wave_convergence_sgd():
    """test/test_model.py:190-199,222-282 on the oracle with TF MomentumOptimizer semantics."""
    net = O.OracleWaveNet(1, [1, 2, 4, 8, 16, 32, 64] * 2, 2, 32, 32, 32, quantization_channels=256, seed=42,
                          faithful=False)
    audio = make_sine_waves().astype(np.float32)
    opt = O.TFOptimizer('sgd', 0.02, 0.95)
    params = net.state_dict()
    initial = None
    for i in range(400):
        loss, _, grads = net.loss_and_grads(audio)
        if initial is None:
            initial = loss
        grads['wavenet/dilated_stack/layer13/dense'] = None   # TF skips variables without gradient (App. A5)
        opt.apply(params, grads)
        net.load_state_dict(params)
    final = float(net.loss(audio))
    assert initial > 0.1 and final < 0.1 and final / initial < 0.02, (initial, final)


def test_choice_from_uniform_matches_numpy_choice():
    rng = np.random.RandomState(7)
    for _ in range(200):
        p = rng.dirichlet(np.ones(256)).astype(np.float32)
        state = rng.get_state()
        expect = rng.choice(np.arange(256), p=p)
        rng.set_state(state)
        u = rng.random_sample()
        assert O.choice_from_uniform(p, u) == expect


def test_tf32_rounding_explains_gradient_deviation():
    """Rounding the GEMM operands to TF32 (what the sm_100a kernels do) keeps logits/loss within
    1e-3 of the exact arithmetic, while gradients at random initialisation move by percents:
    the tolerance used for the end-to-end gradient check on the GPU is a property of the
    arithmetic, not of the kernels."""
    kw = dict(batch_size=1, dilations=[1, 2, 4, 8, 16, 32, 64] * 2, filter_width=2, residual_channels=32,
              dilation_channels=32, quantization_channels=256, skip_channels=32)
    rng = np.random.default_rng(7)
    tt = np.arange(1000) / 16000.
    a = np.clip(0.3 * np.sin(2 * np.pi * 220 * tt)[None] + 0.3 * np.sin(2 * np.pi * 331 * tt)[None] +
                0.1 * rng.standard_normal((1, 1000)), -1, 1).astype(np.float32)
    exact = O.OracleWaveNet(dtype=torch.float64, seed=1, bias_scale=0.1, faithful=False, **kw)
    emul = O.OracleWaveNet(dtype=torch.float64, seed=1, bias_scale=0.1, faithful=False, emulate='tf32', **kw)
    l0, lg0, g0 = exact.loss_and_grads(a)
    l1, lg1, g1 = emul.loss_and_grads(a)
    assert abs(l0 - l1) < 1e-3 * abs(l0)
    assert np.abs(lg0 - lg1).max() < 1e-3 * np.abs(lg0).max()
    k = 'wavenet/causal_layer/filter'
    dev = np.linalg.norm(g0[k] - g1[k]) / np.linalg.norm(g0[k])
    assert 5e-3 < dev < 5e-2, dev
