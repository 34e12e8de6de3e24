"""Parity of every sm_100a kernel, called through the C ABI, against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import wavenet_oracle as O
from wn_helpers import GRAD_RTOL, LOGIT_RTOL, dev, p, rel_err, stream


@pytest.fixture(scope='module')
def lib():
    from wavenet import _lib
    _lib.require_cuda()
    return _lib.load()


# ----------------------------------------------------------------------------- mu-law
def _edge_inputs(q):
    from wavenet.ops import mu_law_tables
    thr, _ = mu_law_tables(q)
    bits = thr.view(np.int32).astype(np.int64)
    near = np.concatenate([bits + k for k in range(-3, 4)]).astype(np.int32).view(np.float32)
    near = near[np.abs(near) <= 1.0]
    special = np.array([-1.0, 1.0, 0.0, -0.0, 1e-30, -1e-30, 0.5, -0.5], np.float32)
    return np.concatenate([near, special])


@pytest.mark.parametrize('q', [256, 128, 123])
def test_mulaw_encode_bit_exact(q):
    import wavenet
    rng = np.random.default_rng(q)
    x = np.concatenate([rng.uniform(-1, 1, 1 << 20).astype(np.float32), _edge_inputs(q),
                        np.linspace(-1, 1, 10001).astype(np.float32),
                        (rng.standard_normal(1 << 16) * 1e-3).astype(np.float32)])
    got = wavenet.mu_law_encode(x, q).cpu().numpy()
    assert got.dtype == np.int32
    np.testing.assert_array_equal(got, O.mu_law_encode(x, q))


def test_mulaw_reference_vectors():
    import wavenet
    x = np.array([-1.0, 1.0, 0.6, -0.25, 0.01, 0.33, -0.9999, 0.42, 0.1, -0.45], np.float32)
    np.testing.assert_array_equal(wavenet.mu_law_encode(x, 256).cpu().numpy(),
                                  np.array([0, 255, 243, 32, 157, 230, 0, 235, 203, 18], np.int32))
    np.random.seed(42)
    x = np.random.uniform(-1, 1, 2048).astype(np.float32)     # test_mu_law.py:126-137
    np.testing.assert_array_equal(wavenet.mu_law_encode(x, 256).cpu().numpy(), O.mu_law_encode(x, 256))
    assert wavenet.mu_law_encode(np.zeros(7, np.float32), 256).cpu().numpy().tolist() == [128] * 7
    assert wavenet.mu_law_encode(np.zeros((0,), np.float32), 256).numel() == 0    # empty input
    assert tuple(wavenet.mu_law_encode(np.zeros((2, 5, 1), np.float32), 256).shape) == (2, 5, 1)


@pytest.mark.parametrize('q', [256, 128])
def test_mulaw_decode_bit_exact(q):
    import wavenet
    ids = np.concatenate([np.arange(q), np.random.default_rng(0).integers(0, q, 5000)]).astype(np.int32)
    got = wavenet.mu_law_decode(ids, q).cpu().numpy()
    np.testing.assert_array_equal(got, O.mu_law_decode(ids, q))
    # decode -> encode round trip of every level (test_mu_law.py:37-51) and exact range (:53-68)
    np.testing.assert_array_equal(wavenet.mu_law_encode(wavenet.mu_law_decode(np.arange(q), q), q).cpu().numpy(),
                                  np.arange(q))
    assert got[:q].max() == 1.0 and got[:q].min() == -1.0


def test_mulaw_full_size_idempotence():
    """BASELINE size (4 x 100000 samples): encode(decode(encode(x))) == encode(x)."""
    import wavenet
    x = torch.rand(4, 100000, device='cuda') * 2 - 1
    e = wavenet.mu_law_encode(x, 256)
    assert int(e.min()) >= 0 and int(e.max()) <= 255
    e2 = wavenet.mu_law_encode(wavenet.mu_law_decode(e, 256), 256)
    assert torch.equal(e, e2)


# ----------------------------------------------------------------------------- causal_conv
def test_causal_conv_reference_tests():
    import wavenet
    x1 = np.arange(1, 21, dtype=np.float32)
    x = np.append(x1, x1).reshape(2, 20, 1)
    f = np.array([1, 1], dtype=np.float32).reshape(2, 1, 1)
    ref = np.convolve(x1, [1, 0, 0, 0, 1])[:-4]
    ref = np.append(ref, ref).reshape(2, 20, 1)
    np.testing.assert_array_equal(wavenet.causal_conv(x, f, 4).cpu().numpy(), ref)       # test_causal_conv.py:11-27
    x = np.arange(1, 11, dtype=np.float32).reshape(1, 10, 1)
    f = np.array([0.0, 1.0], dtype=np.float32).reshape(2, 1, 1)
    out = wavenet.causal_conv(x, f, dilation=2).cpu().numpy()                            # :29-58
    assert out.shape == x.shape
    np.testing.assert_array_equal(out, x)


@pytest.mark.parametrize('shape', [(1, 50, 3, 4, 2, 1), (2, 37, 5, 2, 2, 4), (3, 64, 8, 8, 2, 16),
                                   (1, 20, 2, 3, 2, 32), (2, 41, 4, 3, 3, 2), (1, 100, 1, 8, 4, 1),
                                   (2, 90, 3, 5, 5, 3)])
def test_causal_conv_vs_oracle_reshape_chain(shape):
    import wavenet
    b, t, cin, cout, width, d = shape
    rng = np.random.default_rng(1)
    x = rng.standard_normal((b, t, cin)).astype(np.float32)
    w = rng.standard_normal((width, cin, cout)).astype(np.float32)
    ref = O.causal_conv(torch.tensor(x, dtype=torch.float64), torch.tensor(w, dtype=torch.float64), d).numpy()
    np.testing.assert_allclose(wavenet.causal_conv(x, w, d).cpu().numpy(), ref, atol=1e-5)


def test_time_to_batch_roundtrip():
    import wavenet
    x = torch.arange(2 * 12 * 3, dtype=torch.float32).reshape(2, 12, 3)
    for d in (1, 2, 3, 4, 6):
        a = wavenet.time_to_batch(x, d)
        np.testing.assert_array_equal(a.cpu().numpy(), O.time_to_batch(x, d).numpy())
        np.testing.assert_array_equal(wavenet.batch_to_time(a, d).cpu().numpy(), x.numpy())


# ----------------------------------------------------------------------------- front end
@pytest.mark.parametrize('B,T,Q,R', [(1, 1000, 256, 32), (3, 77, 128, 16), (2, 1, 256, 32)])
def test_frontend_fwd_bwd(lib, B, T, Q, R):
    rng = np.random.default_rng(3)
    ids = rng.integers(0, Q, (B, T)).astype(np.int32)
    ids[0, min(3, T - 1)] = Q + 5          # out-of-range id -> all-zero one-hot row (App. A14)
    wc = rng.standard_normal((2, Q, R)).astype(np.float32)
    d_ids, d_wc = dev(ids, torch.int32), dev(wc)
    x0 = torch.empty(B * T, R, device='cuda')
    assert lib.wn_frontend_fwd(p(d_ids), p(d_wc), p(x0), B, T, Q, R, stream()) == 0
    # oracle: dense one-hot through the causal conv (model.py:227-234, 518-531)
    oh = np.zeros((B, T, Q), np.float64)
    valid = ids < Q
    bb, tt = np.nonzero(valid)
    oh[bb, tt, ids[bb, tt]] = 1.0
    w64 = torch.tensor(wc, dtype=torch.float64, requires_grad=True)
    ref = O.causal_conv(torch.tensor(oh), w64, 1)
    np.testing.assert_allclose(x0.cpu().numpy().reshape(B, T, R), ref.detach().numpy(), atol=1e-6)
    dx0 = rng.standard_normal((B, T, R)).astype(np.float32)
    (ref * torch.tensor(dx0, dtype=torch.float64)).sum().backward()
    gwc = torch.zeros(2, Q, R, device='cuda')
    assert lib.wn_frontend_bwd(p(d_ids), p(dev(dx0)), p(gwc), B, T, Q, R, stream()) == 0
    np.testing.assert_allclose(gwc.cpu().numpy(), w64.grad.numpy(), atol=1e-3 * max(1.0, np.sqrt(B * T / Q)))


# ----------------------------------------------------------------------------- residual block
def _block_oracle(x, wf, wg, wd, prebias, bd, d, is_last, gz, gx):
    """model.py:236-330 in float64 + autograd of sum(z*gz) + sum(x'*gx)."""
    t64 = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    X, WF, WG, WD, PB, BD = t64(x), t64(wf), t64(wg), t64(wd), t64(prebias), t64(bd)
    C_ = x.shape[-1]
    f = O.causal_conv_closed_form(X, WF, d) + PB[:, None, :C_]
    g = O.causal_conv_closed_form(X, WG, d) + PB[:, None, C_:]
    z = torch.tanh(f) * torch.sigmoid(g)
    loss = (z * torch.tensor(gz, dtype=torch.float64)).sum()
    xo = None
    if not is_last:
        xo = X + z @ WD[0] + BD
        loss = loss + (xo * torch.tensor(gx, dtype=torch.float64)).sum()
    loss.backward()
    return dict(z=z.detach().numpy(), xo=None if xo is None else xo.detach().numpy(), dx=X.grad.numpy(),
                dwf=WF.grad.numpy(), dwg=WG.grad.numpy(), dwd=None if is_last else WD.grad.numpy(),
                dpb=PB.grad.numpy(), dbd=None if is_last else BD.grad.numpy())


@pytest.mark.parametrize('C_,B,T,d,is_last', [(32, 1, 1000, 1, 0), (32, 1, 1000, 64, 0), (32, 3, 333, 8, 0),
                                              (32, 2, 700, 512, 0), (32, 1, 999, 4, 1), (16, 2, 515, 16, 0),
                                              (16, 1, 64, 256, 1), (32, 1, 16, 2, 0), (32, 2, 5, 1, 0),
                                              (32, 1, 100, 128, 0), (32, 3, 640, 512, 0), (32, 1, 257, 256, 1),
                                              (32, 2, 1153, 127, 0), (16, 3, 300, 512, 0),
                                              # wide blocks (block_wide16.cu: two-tap GEMM, gate / dpre in the epilogues, 16-bit storage)
                                              (128, 1, 1000, 1, 0), (128, 2, 700, 512, 0), (128, 3, 333, 8, 0),
                                              (128, 1, 999, 4, 1), (128, 2, 130, 256, 0), (64, 2, 515, 16, 0),
                                              (64, 1, 300, 2, 1), (192, 2, 257, 64, 0)])
def test_block_fwd_bwd(lib, C_, B, T, d, is_last):
    """wn_block_fwd / wn_block_bwd: for 32 channels these run the PRODUCTION kernels of the training step
    (block_fwd_chain_kernel; block_bwd_chain_f + reduce) as a one-layer network, for multiples of 64 channels the wide
    blocks in 16-bit storage.  Shapes cover T not a multiple of the 128-step tile, d >= T (no valid past tap), d = 512
    with B = 3, 16 channels.  Tolerances of the 16-bit path: every stored activation / gradient is rounded to fp16 (2^-11)."""
    wide = C_ % 64 == 0
    tol_z, tol_x, tol_g = (3e-3, 2e-3, 8e-3) if wide else (6e-4, 2e-5, GRAD_RTOL)
    rng = np.random.default_rng(C_ + T + d)
    M = B * T
    lim = np.sqrt(6.0 / (4 * C_))
    x = rng.standard_normal((B, T, C_)).astype(np.float32)
    wf = rng.uniform(-lim, lim, (2, C_, C_)).astype(np.float32)
    wg = rng.uniform(-lim, lim, (2, C_, C_)).astype(np.float32)
    wd = rng.uniform(-lim, lim, (1, C_, C_)).astype(np.float32)
    prebias = (0.2 * rng.standard_normal((B, 2 * C_))).astype(np.float32)
    bd = (0.2 * rng.standard_normal(C_)).astype(np.float32)
    gz = rng.standard_normal((B, T, C_)).astype(np.float32)
    gx = rng.standard_normal((B, T, C_)).astype(np.float32)
    ref = _block_oracle(x, wf, wg, wd, prebias, bd, d, is_last, gz, gx)

    ldz = 3 * C_                      # z lives in column block 1 of a wider Zcat
    dx_, dwf_, dwg_, dwd_, dpb_, dbd_ = dev(x), dev(wf), dev(wg), dev(wd), dev(prebias), dev(bd)
    zc = torch.zeros(M, ldz, device='cuda')
    xo = torch.zeros(M, C_, device='cuda')
    rc = lib.wn_block_fwd(p(dx_), p(xo), C.c_void_p(zc.data_ptr() + 4 * C_), ldz, p(dwf_), p(dwg_), p(dwd_),
                          p(dpb_), p(dbd_), B, T, d, C_, is_last, stream())
    assert rc == 0
    z = zc[:, C_:2 * C_].cpu().numpy().reshape(B, T, C_)
    # forward block = split-precision (3xTF32) products; only the stored z is tf32-rounded (2^-11)
    assert rel_err(z, ref['z']) < tol_z
    assert float(zc[:, :C_].abs().max()) == 0.0 and float(zc[:, 2 * C_:].abs().max()) == 0.0
    if not is_last:
        assert rel_err(xo.cpu().numpy().reshape(B, T, C_), ref['xo']) < tol_x

    dzs = torch.zeros(M, ldz, device='cuda')
    dzs[:, C_:2 * C_] = dev(gz).reshape(M, C_)
    dxo = dev(gx).reshape(M, C_)
    dx = torch.zeros(M, C_, device='cuda')
    dpre = torch.zeros(M, 2 * C_, device='cuda')
    gwf, gwg, gwd = torch.zeros(2, C_, C_, device='cuda'), torch.zeros(2, C_, C_, device='cuda'), torch.zeros(C_, C_, device='cuda')
    gpb, gbd = torch.zeros(B, 2 * C_, device='cuda'), torch.zeros(C_, device='cuda')
    rc = lib.wn_block_bwd(p(dx_), p(dxo), C.c_void_p(dzs.data_ptr() + 4 * C_), ldz, p(dx), p(dpre),
                          C.c_void_p(zc.data_ptr() + 4 * C_), p(dwf_), p(dwg_), p(dwd_), p(dpb_), p(gwf), p(gwg),
                          p(gwd), p(gpb), p(gbd), B, T, d, C_, is_last, stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert rel_err(dx.cpu().numpy().reshape(B, T, C_), ref['dx']) < tol_g
    assert rel_err(gwf.cpu().numpy(), ref['dwf']) < tol_g
    assert rel_err(gwg.cpu().numpy(), ref['dwg']) < tol_g
    assert rel_err(gpb.cpu().numpy(), ref['dpb']) < tol_g
    if not is_last:
        assert rel_err(gwd.cpu().numpy(), ref['dwd'][0]) < tol_g
        assert rel_err(gbd.cpu().numpy(), ref['dbd']) < tol_g


# ----------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize('mode,M,N,K', [(0, 1000, 512, 1600), (0, 333, 32, 448), (0, 4096, 256, 512),
                                        (1, 1000, 512, 256), (1, 517, 1600, 512), (1, 100, 32, 32),
                                        (2, 1600, 512, 3000), (2, 32, 256, 1000), (2, 512, 512, 777)])
def test_gemm_tf32(lib, mode, M, N, K):
    rng = np.random.default_rng(M + N + K)
    if mode == 0:
        a, b = rng.standard_normal((M, K)), rng.standard_normal((K, N))
        ref = a @ b
    elif mode == 1:
        a, b = rng.standard_normal((M, K)), rng.standard_normal((N, K))
        ref = a @ b.T
    else:
        a, b = rng.standard_normal((K, M)), rng.standard_normal((K, N))
        ref = a.T @ b
    da, db = dev(a), dev(b)
    c = torch.zeros(M, N, device='cuda')
    rc = lib.wn_gemm_tf32(mode, p(da), da.shape[1], p(db), db.shape[1], p(c), N, M, N, K, None, None, 0, 0,
                          7 if mode == 2 else 1, stream())
    assert rc == 0
    assert rel_err(c.cpu().numpy(), ref) < LOGIT_RTOL


def test_gemm_epilogues(lib):
    rng = np.random.default_rng(5)
    M, N, K = 300, 64, 96
    a, b, bias = rng.standard_normal((M, K)), rng.standard_normal((K, N)), rng.standard_normal(N)
    mask = rng.standard_normal((M, N))
    da, db, dbias, dmask = dev(a), dev(b), dev(bias), dev(mask)
    c = torch.zeros(M, N, device='cuda')
    assert lib.wn_gemm_tf32(0, p(da), K, p(db), N, p(c), N, M, N, K, p(dbias), None, 0, 1 | 2, 1, stream()) == 0
    assert rel_err(c.cpu().numpy(), np.maximum(a @ b + bias, 0)) < LOGIT_RTOL
    assert lib.wn_gemm_tf32(0, p(da), K, p(db), N, p(c), N, M, N, K, None, p(dmask), N, 2, 1, stream()) == 0
    assert rel_err(c.cpu().numpy(), (a @ b) * (mask > 0)) < LOGIT_RTOL


@pytest.mark.parametrize('M,N,K,split', [(1000, 512, 1600, 0), (333, 32, 32, 0), (4096, 256, 512, 0), (517, 1600, 512, 0),
                                         (1600, 512, 3001, 9), (32, 256, 1000, 4), (512, 512, 777, 3), (130, 132, 40, 0)])
def test_gemm_nt_umma(lib, M, N, K, split):
    """tcgen05 GEMM: C = A[M,K] . B[N,K]^T (+ transposed copy, split-K atomics) vs float64."""
    rng = np.random.default_rng(M + N + K)
    # kind::tf32 consumes the upper 19 bits of each fp32 operand: callers pre-round (cvt.rna), and so
    # does the test -- products are then exact and only the fp32 accumulation order differs
    a = O.round_tf32(torch.tensor(rng.standard_normal((M, K)))).numpy()
    b = O.round_tf32(torch.tensor(rng.standard_normal((N, K)))).numpy()
    ref = a @ b.T
    lda = (K + 3) // 4 * 4
    da = torch.zeros(M, lda, device='cuda'); da[:, :K] = dev(a)
    db = torch.zeros(N, lda, device='cuda'); db[:, :K] = dev(b)
    c = torch.zeros(M, N, device='cuda')
    ldct = (M + 3) // 4 * 4
    ct = torch.zeros(N, ldct, device='cuda')
    flags = 4 if split else 0
    rc = lib.wn_gemm_nt_umma(p(da), lda, p(db), lda, p(c), N, None if split else p(ct), ldct, M, N, K, None, None, 0,
                             flags, split, stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert rel_err(c.cpu().numpy(), ref) < 1e-5      # fp32 accumulation order only
    if not split:
        np.testing.assert_array_equal(ct[:, :M].cpu().numpy(), c.cpu().numpy().T)


@pytest.mark.parametrize('mode,M,N,K,split', [(0, 1000, 512, 1600, 0), (0, 333, 32, 448, 0), (0, 4096, 256, 512, 0),
                                              (0, 130, 64, 40, 0), (1, 517, 1600, 512, 0),
                                              (2, 1600, 512, 3001, 9), (2, 32, 256, 1000, 4), (2, 512, 512, 777, 3),
                                              (2, 64, 32, 100000, 148)])
def test_gemm_umma_modes(lib, mode, M, N, K, split):
    """tcgen05 GEMM with MN-major operands (32-byte-atom swizzle): NN / NT / TN vs float64 on tf32-exact inputs."""
    rng = np.random.default_rng(mode + M + N + K)
    r = lambda *s: O.round_tf32(torch.tensor(rng.standard_normal(s))).numpy()
    if mode == 0:
        a, b = r(M, K), r(K, N)
        ref = a @ b
    elif mode == 1:
        a, b = r(M, K), r(N, K)
        ref = a @ b.T
    else:
        a, b = r(K, M), r(K, N)
        ref = a.T @ b
    def padded(x):
        ld = (x.shape[1] + 3) // 4 * 4
        t = torch.zeros(x.shape[0], ld, device='cuda')
        t[:, :x.shape[1]] = dev(x)
        return t, ld
    da, lda = padded(a)
    db, ldb = padded(b)
    c = torch.zeros(M, N, device='cuda')
    rc = lib.wn_gemm_umma(mode, p(da), lda, p(db), ldb, p(c), N, M, N, K, None, None, 0, 0, split, stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert rel_err(c.cpu().numpy(), ref) < 1e-5      # fp32 accumulation order only


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (1000, 512, 1600), (333, 256, 512), (257, 96, 200), (4100, 512, 512)])
def test_gemm_f16_nt(lib, M, N, K):
    """fp16-operand tcgen05 GEMM of the forward chain: exact products of fp16 inputs, fp32 accumulate; fp32 output
    (tf32-rounded, relu, bias) and its fp16 copy."""
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K)).astype(np.float16)
    b = (0.1 * rng.standard_normal((N, K))).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float32)
    da, db = torch.tensor(a, device='cuda'), torch.tensor(b, device='cuda')
    c = torch.full((M, N), -7.0, device='cuda')
    c16 = torch.full((M, N), -7.0, device='cuda', dtype=torch.float16)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    assert lib.wn_gemm_f16_nt(p(da), K, p(db), K, p(c), N, None, 0, M, N, K, None, None, 0, 1.0, 0, stream()) == 0
    torch.cuda.synchronize()
    assert rel_err(c.cpu().numpy(), ref) < 1e-5
    assert lib.wn_gemm_f16_nt(p(da), K, p(db), K, p(c), N, p(c16), N, M, N, K, p(dev(bias)), None, 0, 1.0, 1 | 2, stream()) == 0
    torch.cuda.synchronize()
    ref = np.maximum(ref + bias, 0)
    out = c.cpu().numpy()
    assert rel_err(out, ref) < LOGIT_RTOL
    np.testing.assert_array_equal(out, O.round_tf32(torch.tensor(out)).numpy())          # tf32-exact values ...
    np.testing.assert_allclose(c16.cpu().numpy().astype(np.float32), out, rtol=1e-3, atol=1e-7)   # the fp16 copy (one rounding)
    # gradient-chain form: relu mask from another matrix, fp32 copy scaled out of the fp16 domain
    mask = rng.standard_normal((M, N)).astype(np.float32)
    assert lib.wn_gemm_f16_nt(p(da), K, p(db), K, p(c), N, p(c16), N, M, N, K, None, p(dev(mask)), N, 0.25, 2, stream()) == 0
    torch.cuda.synchronize()
    ref = (a.astype(np.float64) @ b.astype(np.float64).T) * (mask > 0)
    assert rel_err(c.cpu().numpy(), 0.25 * ref) < LOGIT_RTOL
    assert rel_err(c16.cpu().numpy().astype(np.float64), ref) < 1e-3
    # input-gradient form of the training step: fp16 copy only, bias gradient (column sums) collected by the epilogue
    cs = torch.full((N,), 3.0, device='cuda')
    c16.fill_(-7.0)
    assert lib.wn_gemm_f16_nt_colsum(p(da), K, p(db), K, None, 0, p(c16), N, M, N, K, 1.0, p(cs), 0.5, stream()) == 0
    torch.cuda.synchronize()
    got16 = c16.cpu().numpy().astype(np.float64)
    assert rel_err(got16, a.astype(np.float64) @ b.astype(np.float64).T) < 1e-3
    np.testing.assert_allclose(cs.cpu().numpy() - 3.0, 0.5 * got16.sum(0), rtol=2e-4, atol=2e-4 * np.abs(got16).sum(0).max())


@pytest.mark.parametrize('M,N,K,split', [(128, 256, 64, 1), (512, 256, 5000, 6), (1600, 512, 3333, 3), (64, 64, 200, 2),
                                          (1024, 256, 2000, 1), (1088, 512, 700, 4)])      # >= 1024 rows: 256-row tiles
def test_gemm_f16_tn(lib, M, N, K, split):
    """Weight-gradient form on fp16 MN-major operands (both read as they lie in memory), split-K atomics, scaled."""
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((K, M)).astype(np.float16)
    b = rng.standard_normal((K, N)).astype(np.float16)
    da, db = torch.tensor(a, device='cuda'), torch.tensor(b, device='cuda')
    c0 = rng.standard_normal((M, N)).astype(np.float32)
    c = dev(c0)
    assert lib.wn_gemm_f16_tn(p(da), M, p(db), N, p(c), N, M, N, K, 0.5, split, stream()) == 0
    torch.cuda.synchronize()
    ref = c0 + 0.5 * (a.astype(np.float64).T @ b.astype(np.float64))
    assert rel_err(c.cpu().numpy(), ref) < 1e-5
    assert lib.wn_gemm_f16_tn(p(da), M, p(db), N, p(c), N, M + 32, N, K, 1.0, 1, stream()) == -3


def test_gemm_f16_rejects_unaligned(lib):
    a = torch.zeros(64, 44, device='cuda', dtype=torch.float16)
    c = torch.zeros(64, 64, device='cuda')
    assert lib.wn_gemm_f16_nt(p(a), 44, p(a), 44, p(c), 64, None, 0, 64, 64, 44, None, None, 0, 1.0, 0, stream()) == -3


def test_gemm_umma_rejects_unaligned_mn(lib):
    a = torch.zeros(64, 48, device='cuda')
    c = torch.zeros(64, 48, device='cuda')
    assert lib.wn_gemm_umma(0, p(a), 48, p(a), 48, p(c), 48, 64, 48, 48, None, None, 0, 0, 0, stream()) == -3


def test_gemm_nt_umma_epilogues(lib):
    rng = np.random.default_rng(5)
    M, N, K = 300, 200, 96
    a = O.round_tf32(torch.tensor(rng.standard_normal((M, K)))).numpy()
    b = O.round_tf32(torch.tensor(rng.standard_normal((N, K)))).numpy()
    bias = rng.standard_normal(N)
    mask = rng.standard_normal((M, N))
    da, db, dbias, dmask = dev(a), dev(b), dev(bias), dev(mask)
    c = torch.zeros(M, N, device='cuda')
    assert lib.wn_gemm_nt_umma(p(da), K, p(db), K, p(c), N, None, 0, M, N, K, p(dbias), None, 0, 1 | 2, 1, stream()) == 0
    assert rel_err(c.cpu().numpy(), np.maximum(a @ b.T + bias, 0)) < LOGIT_RTOL
    assert lib.wn_gemm_nt_umma(p(da), K, p(db), K, p(c), N, None, 0, M, N, K, None, p(dmask), N, 2, 1, stream()) == 0
    assert rel_err(c.cpu().numpy(), (a @ b.T) * (mask > 0)) < LOGIT_RTOL


# ----------------------------------------------------------------------------- softmax cross entropy
@pytest.mark.parametrize('B,T,Q', [(1, 1000, 256), (3, 50, 128), (2, 7, 512), (1, 1, 256)])
def test_softmax_xent(lib, B, T, Q):
    rng = np.random.default_rng(Q + T)
    M = B * T
    logits = (3 * rng.standard_normal((M, Q))).astype(np.float32)
    ids = rng.integers(0, Q, (B, T)).astype(np.int32)
    L64 = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    oh = torch.nn.functional.one_hot(torch.tensor(ids, dtype=torch.int64), Q).to(torch.float64)
    shifted = torch.nn.functional.pad(oh[:, 1:, :], (0, 0, 0, 1)).reshape(M, Q)
    loss_ref = O._TFSoftmaxXent.apply(L64, shifted).mean()
    loss_ref.backward()
    dl = dev(logits)
    partials = torch.zeros(4096, device='cuda')
    out = torch.zeros((), device='cuda')
    assert lib.wn_softmax_xent(p(dl), p(dev(ids, torch.int32)), B, T, Q, p(partials), 4096, p(out), 1, stream()) == 0
    assert abs(float(out) - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert rel_err(dl.cpu().numpy(), L64.grad.numpy()) < 1e-3      # gradient is stored tf32-rounded


@pytest.mark.parametrize('B,T,K,with_bias', [(1, 1000, 512, True), (3, 77, 128, True), (2, 4099, 64, False), (1, 1, 512, True),
                                             (1, 40000, 512, True)])
def test_post2_xent_fused(lib, B, T, K, with_bias):
    """postprocess2 GEMM + TF softmax cross entropy + fp16 gradient + bias column sums in one kernel (csrc/post_xent.cu)
    against the oracle's xent on the logits of the same fp16-rounded operands: loss 1e-5, gradient 1e-3 (it is fp16),
    bias gradient 1e-3.  model.py:438-440, 654-666."""
    Q = 256
    rng = np.random.default_rng(K + T)
    M = B * T
    a16 = torch.tensor(np.maximum(rng.standard_normal((M, K)), 0).astype(np.float32)).half()
    w16 = torch.tensor((rng.standard_normal((Q, K)) * (2.0 / np.sqrt(K))).astype(np.float32)).half()
    bias = (0.5 * rng.standard_normal(Q)).astype(np.float32) if with_bias else None
    ids = rng.integers(0, Q, (B, T)).astype(np.int32)
    logits = a16.double() @ w16.double().T + (torch.tensor(bias, dtype=torch.float64) if with_bias else 0.0)
    L64 = logits.clone().requires_grad_(True)
    oh = torch.nn.functional.one_hot(torch.tensor(ids, dtype=torch.int64), Q).to(torch.float64)
    shifted = torch.nn.functional.pad(oh[:, 1:, :], (0, 0, 0, 1)).reshape(M, Q)
    loss_ref = O._TFSoftmaxXent.apply(L64, shifted).mean()
    loss_ref.backward()
    grad_ref = L64.grad.numpy() * M                      # (softmax - onehot), unscaled
    da, dw = a16.cuda().contiguous(), w16.cuda().contiguous()
    partials = torch.zeros(4096, device='cuda')
    out = torch.zeros((), device='cuda')
    g16 = torch.zeros(M, Q, dtype=torch.float16, device='cuda')
    bgrad = torch.full((Q,), 0.25, device='cuda')
    gs = 1.5
    dbias, dids = (dev(bias) if with_bias else None), dev(ids, torch.int32)      # (kept alive across the launch)
    rc = lib.wn_post2_xent(p(da), K, p(dw), K, p(dbias) if with_bias else None, p(dids), B, T, K, Q,
                           p(partials), p(out), p(g16), gs, p(bgrad), 1.0 / gs, stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert abs(float(out) - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert rel_err(g16.float().cpu().numpy() / gs, grad_ref) < 1e-3
    assert rel_err(bgrad.cpu().numpy() - 0.25, grad_ref.sum(0)) < 2e-3      # sums of the fp16-rounded gradient
    assert lib.wn_post2_xent(p(da), K, p(dw), K, None, p(dids), B, T, K, 128, p(partials), p(out), p(g16), gs,
                             None, 0.0, stream()) == -2


# ----------------------------------------------------------------------------- optimizers
@pytest.mark.parametrize('kind', ['adam', 'sgd', 'rmsprop'])
def test_optimizers_bit_exact(kind):
    from wavenet.ops import optimizer_factory
    rng = np.random.default_rng(11)
    n = 10007
    w0 = rng.standard_normal(n).astype(np.float32)
    ref = O.TFOptimizer(kind, 1e-3 if kind != 'sgd' else 0.02, 0.95)
    params = {'w': w0.copy()}
    opt = optimizer_factory[kind](learning_rate=ref.lr, momentum=0.95)
    w = dev(w0)
    for _ in range(4):
        g = (rng.standard_normal(n) * 0.1).astype(np.float32)
        ref.apply(params, {'w': g})
        opt.apply(w, dev(g))
    got = w.cpu().numpy()
    np.testing.assert_allclose(got, params['w'], rtol=0, atol=2e-7)
    assert np.mean(got == params['w']) > 0.99


# ----------------------------------------------------------------------------- sampler
def test_sampler_bit_exact(lib):
    rng = np.random.RandomState(3)
    rows, Q = 4000, 256
    pr = rng.dirichlet(np.ones(Q) * 0.3, size=rows).astype(np.float32)
    pr[5, 10:200] = 0.0
    pr[5] /= pr[5].sum()
    u = rng.random_sample(rows)
    u[:4] = [0.0, 1.0 - 2 ** -53, 0.5, 1e-300]
    out = torch.zeros(rows, dtype=torch.int32, device='cuda')
    assert lib.wn_sample(p(dev(pr)), p(dev(u, torch.float64)), rows, Q, p(out), stream()) == 0
    ref = np.array([O.choice_from_uniform(pr[i], u[i]) for i in range(rows)])
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    # and np.random.choice itself on a replayed generator state
    r2 = np.random.RandomState(9)
    state = r2.get_state()
    expect = [r2.choice(np.arange(Q), p=pr[i]) for i in range(100)]
    r2.set_state(state)
    u2 = r2.random_sample(100)
    out2 = torch.zeros(100, dtype=torch.int32, device='cuda')
    assert lib.wn_sample(p(dev(pr[:100])), p(dev(u2, torch.float64)), 100, Q, p(out2), stream()) == 0
    np.testing.assert_array_equal(out2.cpu().numpy(), np.array(expect))
