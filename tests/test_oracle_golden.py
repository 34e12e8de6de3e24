"""Pins the CPU oracle against every golden vector / exact equality the reference's own tests
hold for the hot path (test/test_mu_law.py, test/test_causal_conv.py)."""
import numpy as np
import torch

import wavenet_oracle as O

QUANT_LEVELS = 256


def manual_mu_law_encode(signal, quantization_channels):
    """test/test_mu_law.py:11-22 evaluated in float32 (what it was under the numpy of the
    reference's era; numpy >= 2 would silently promote to float64)."""
    f32 = np.float32
    mu = f32(quantization_channels - 1)
    signal = np.asarray(signal, dtype=np.float32)
    magnitude = np.log(f32(1) + mu * np.abs(signal)) / np.log(f32(1.) + mu)
    signal = np.sign(signal) * magnitude
    signal = (signal + f32(1)) / f32(2) * mu + f32(0.5)
    assert signal.dtype == np.float32
    return signal.astype(np.int32)


def manual_mu_law_decode(signal, quantization_channels):
    """test/test_mu_law.py:25-32 in float32."""
    f32 = np.float32
    mu = quantization_channels - 1
    y = signal.astype(np.float32)
    y = f32(2) * (y / f32(mu)) - f32(1)
    x = np.sign(y) * f32(1.0 / mu) * (np.power(f32(1.0 + mu), np.abs(y)) - f32(1.0))
    assert x.dtype == np.float32
    return x


def test_encode_precomputed():
    # test_mu_law.py:113-124 known-answer vector
    x = np.array([-1.0, 1.0, 0.6, -0.25, 0.01, 0.33, -0.9999, 0.42, 0.1, -0.45]).astype(np.float32)
    expected = np.array([0, 255, 243, 32, 157, 230, 0, 235, 203, 18]).astype(np.int32)
    np.testing.assert_array_equal(O.mu_law_encode(x, 256), expected)


def test_decode_encode_roundtrip_all_levels():
    # test_mu_law.py:37-51
    x = np.arange(QUANT_LEVELS)
    np.testing.assert_array_equal(O.mu_law_encode(O.mu_law_decode(x, QUANT_LEVELS), QUANT_LEVELS), x)


def test_min_max_range():
    # test_mu_law.py:53-68
    d = O.mu_law_decode(np.arange(QUANT_LEVELS), QUANT_LEVELS)
    assert abs(d.max() - 1.0) < 1e-10 and abs(d.min() + 1.0) < 1e-10


def test_encode_decode_shift():
    # test_mu_law.py:70-85
    x = np.linspace(-1, 1, 1000).astype(np.float32)
    rt = O.mu_law_decode(O.mu_law_encode(x, QUANT_LEVELS), QUANT_LEVELS)
    slope, icpt = np.polyfit(x, rt, 1)
    assert abs(slope - 1.0) < 1e-4 and abs(icpt) < 1e-4


def test_encode_decode():
    # test_mu_law.py:87-104
    x = np.linspace(-1, 1, 1000).astype(np.float32)
    x1 = O.mu_law_decode(O.mu_law_encode(x, 256), 256)
    np.testing.assert_allclose(x, x1, rtol=1e-1, atol=0.05)
    x2 = O.mu_law_decode(O.mu_law_encode(x1, 256), 256)
    np.testing.assert_allclose(x1, x2)


def test_encode_is_surjective():
    # test_mu_law.py:106-111
    x = np.linspace(-1, 1, 10000).astype(np.float32)
    assert len(np.unique(O.mu_law_encode(x, 123))) == 123


def test_encode_seeded_equalities():
    # test_mu_law.py:126-178
    np.random.seed(42)
    x = np.random.uniform(-1, 1, 2048).astype(np.float32)
    np.testing.assert_array_equal(manual_mu_law_encode(x, 256), O.mu_law_encode(x, 256))
    np.random.seed(1944)
    x = np.zeros(1024).astype(np.float32)
    x.fill(np.random.uniform(-1, 1))
    np.testing.assert_array_equal(manual_mu_law_encode(x, 256), O.mu_law_encode(x, 256))
    x = np.arange(-1.0, 1.0, 2.0 / 1024).astype(np.float32)
    np.testing.assert_array_equal(manual_mu_law_encode(x, 256), O.mu_law_encode(x, 256))
    x = np.zeros(1024).astype(np.float32)
    np.testing.assert_array_equal(manual_mu_law_encode(x, 256), O.mu_law_encode(x, 256))
    assert O.mu_law_encode(x, 256)[0] == 128


def test_decode_seeded_equalities():
    # test_mu_law.py:205-261 (channels = 128, seed 40)
    np.random.seed(40)
    for x in (np.random.uniform(-1, 1, 512), np.full(512, np.random.uniform(-1, 1)),
              np.arange(-1.0, 1.0, 2.0 / 512), np.zeros(100)):
        y = manual_mu_law_encode(x, 128)
        np.testing.assert_array_equal(manual_mu_law_decode(y, 128), O.mu_law_decode(y, 128))


def test_causal_conv_golden():
    # test_causal_conv.py:11-27
    x1 = np.arange(1, 21, dtype=np.float32)
    x = np.append(x1, x1).reshape(2, 20, 1)
    f = np.array([1, 1], dtype=np.float32).reshape(2, 1, 1)
    out = O.causal_conv(torch.tensor(x), torch.tensor(f), 4).numpy()
    ref = np.convolve(x1, [1, 0, 0, 0, 1])[:-4]
    ref = np.append(ref, ref).reshape(2, 20, 1)
    np.testing.assert_array_equal(out, ref)


def test_causal_conv_no_time_shift():
    # test_causal_conv.py:29-58
    x = np.arange(1, 11, dtype=np.float32).reshape(1, 10, 1)
    f = np.array([0.0, 1.0], dtype=np.float32).reshape(2, 1, 1)
    out = O.causal_conv(torch.tensor(x), torch.tensor(f), 2).numpy()
    assert out.shape == x.shape
    np.testing.assert_array_equal(out, x)


def test_closed_form_equals_reshape_chain():
    rng = np.random.default_rng(0)
    for (b, t, cin, cout, d) in [(1, 50, 3, 4, 1), (2, 37, 5, 2, 4), (3, 64, 8, 8, 16), (1, 20, 2, 3, 32)]:
        x = torch.tensor(rng.standard_normal((b, t, cin)))
        w = torch.tensor(rng.standard_normal((2, cin, cout)))
        np.testing.assert_allclose(O.causal_conv(x, w, d).numpy(), O.causal_conv_closed_form(x, w, d).numpy(),
                                   atol=1e-12)


def test_time_to_batch_roundtrip():
    x = torch.arange(2 * 12 * 3, dtype=torch.float32).reshape(2, 12, 3)
    for d in (1, 2, 3, 4, 6):
        np.testing.assert_array_equal(O.batch_to_time(O.time_to_batch(x, d), d).numpy(), x.numpy())
