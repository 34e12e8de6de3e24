"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol that
include/wavenet_b200.h declares (no compute without a GPU), the flat parameter layout, the mu-law
tables that make the integer kernels bit exact, and loud failure without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import wavenet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'wavenet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(wn_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from wavenet import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, 'ctypes signature missing for ' + n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.wn_abi_version() == 2


def test_param_layout_matches_reference_shapes():
    from wavenet import _lib
    dil = [2 ** i for i in range(10)] * 5
    cfg = _lib.make_config(dil, 32, 32, 512, 256, None, None, True, False)
    lo = _lib.param_layout(cfg)
    L, R, D, S, Q = 50, 32, 32, 512, 256
    n_params = 2 * Q * R + L * (2 * 2 * R * D + D * R + D * S + 2 * D + R + S) + S * S + S * Q + S + Q
    assert n_params == 1515968                      # SURVEY App. B parameter count for the default net
    offs = [getattr(lo, f) for f in _lib.LAYOUT_FIELDS[:-1] if getattr(lo, f) >= 0]
    assert all(o % 64 == 0 for o in offs) and len(set(offs)) == len(offs)
    assert n_params <= lo.total < n_params + 64 * len(offs)
    cfg = _lib.make_config(dil, 32, 32, 512, 256, 32, 377, True, False)
    assert _lib.param_layout(cfg).gc_embedding >= 0
    assert 1630432 <= _lib.param_layout(cfg).total < 1630432 + 64 * 20   # cfg3 count


def test_unsupported_configs_fail_loudly():
    from wavenet import _lib
    with pytest.raises(NotImplementedError):
        _lib.param_layout(_lib.make_config([1, 2], 24, 24, 32, 256, None, None, False, False))   # R does not divide 256
    with pytest.raises(NotImplementedError):
        _lib.param_layout(_lib.make_config([1, 2], 32, 18, 32, 256, None, None, False, False))   # D not a multiple of 4
    # any other widths are served by the GEMM-built blocks (block_generic.cu)
    assert _lib.param_layout(_lib.make_config([1, 2], 32, 16, 32, 256, None, None, False, False)).total > 0   # R != D
    assert _lib.param_layout(_lib.make_config([1, 2], 128, 128, 512, 256, None, None, True, False)).total > 0
    with pytest.raises(ValueError):
        _lib.make_config(list(range(1, 200)), 32, 32, 32, 256, None, None, False, False)


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    from wavenet import _lib
    lib = _lib.load()
    cfg = _lib.make_config([1, 2, 4], 32, 32, 64, 256, None, None, True, False)
    assert lib.wn_train_workspace_bytes(C.byref(cfg), 0, 100) < 0
    assert lib.wn_train_workspace_bytes(C.byref(cfg), 1, 1000) > 0
    assert lib.wn_forward_workspace_bytes(C.byref(cfg), 2, 1000) < lib.wn_train_workspace_bytes(C.byref(cfg), 2, 1000)
    assert lib.wn_gen_state_bytes(C.byref(cfg), 3) > 3 * 7 * 32 * 4
    assert lib.wn_loss_grad(C.byref(cfg), None, None, None, 0, None, None, None, 1, 10, None, None) < 0
    assert lib.wn_mulaw_encode(None, 5, None, 256, None, None) < 0
    assert lib.wn_block_fwd(None, None, None, 32, None, None, None, None, None, 1, 1, 1, 32, 0, None) < 0
    assert lib.wn_gemm_tf32(0, None, 4, None, 4, None, 4, 4, 4, 4, None, None, 0, 0, 1, None) < 0


@pytest.mark.parametrize('q', [256, 128, 123, 2, 1024])
def test_mulaw_tables_reproduce_the_oracle(q):
    from wavenet.ops import mu_law_tables
    thr, lut = mu_law_tables(q)
    assert thr.dtype == np.float32 and thr.shape == (q - 1,) and np.all(np.diff(thr) > 0)
    np.testing.assert_array_equal(lut, O.mu_law_decode(np.arange(q), q))
    rng = np.random.default_rng(q)
    bits = thr.view(np.int32).astype(np.int64)
    near = np.concatenate([bits + k for k in range(-4, 5)]).astype(np.int32).view(np.float32)
    x = np.concatenate([rng.uniform(-1, 1, 300000).astype(np.float32), near[np.abs(near) <= 1],
                        np.array([-1, 1, 0, -0.0], np.float32)])
    np.testing.assert_array_equal(np.searchsorted(thr, x, side='right'), O.mu_law_encode(x, q))
    # the threshold is the FIRST float32 of its bin
    below = np.nextafter(thr, np.float32(-np.inf))
    np.testing.assert_array_equal(O.mu_law_encode(thr, q), np.arange(1, q))
    np.testing.assert_array_equal(O.mu_law_encode(below, q), np.arange(0, q - 1))


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a machine without CUDA')
def test_product_path_refuses_to_run_without_cuda():
    import wavenet
    from wavenet._lib import WavenetCudaError
    with pytest.raises(WavenetCudaError):
        wavenet.WaveNetModel(1, [1, 2], 2, 32, 32, 32)
    with pytest.raises(WavenetCudaError):
        wavenet.mu_law_encode(np.zeros(4, np.float32), 256)
    with pytest.raises(WavenetCudaError):
        wavenet.causal_conv(np.zeros((1, 4, 1), np.float32), np.zeros((2, 1, 1), np.float32), 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'tensorflow-wavenet_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'wavenet_oracle' not in text and 'oracle/' not in text, os.path.join(dirpath, f)


def test_golden_fixture_matches_oracle():
    """tests/golden/oracle_small.npz (written by tests/golden/make_golden.py) still equals what the
    oracle computes: guards the checker itself against drift."""
    path = os.path.join(ROOT, 'tests', 'golden', 'oracle_small.npz')
    g = np.load(path)
    import json
    kw = json.loads(str(g['config']))
    net = O.OracleWaveNet(seed=int(g['seed']), bias_scale=0.1, faithful=True, **kw)
    loss, logits, grads = net.loss_and_grads(g['audio'], g['gc'] if g['gc'].size else None)
    assert abs(loss - float(g['loss'])) < 1e-5
    np.testing.assert_allclose(logits, g['logits'], atol=2e-5)
    np.testing.assert_allclose(grads['wavenet/postprocessing/postprocess2'], g['grad_post2'], atol=1e-6)
    np.testing.assert_array_equal(O.mu_law_encode(g['mulaw_x'], 256), g['mulaw_ids'])


# ----------------------------------------------------------------------------- data parallel host logic (N > 1)
def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from wavenet.train_step import allreduce_gradients, shard_streams
        # per-rank gradient of a per-rank batch: the all-reduced, rescaled buffer is the global-batch mean
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        scale = allreduce_gradients(g)
        mean = g * scale
        lo, hi = shard_streams(257, rank, world)
        out[rank] = (float(scale), mean[:5].tolist(), float(mean.sum()), lo, hi)
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_and_stream_sharding_world2():
    """World size 2 over gloo on CPU: the same helpers TrainStep / bench.py use over NCCL."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    ref = torch.arange(1000, dtype=torch.float32) * 1.5         # mean of 1x and 2x
    for rank in range(world):
        scale, head, total, lo, hi = res[rank]
        assert scale == 0.5
        assert head == ref[:5].tolist()
        assert abs(total - float(ref.sum())) < 1e-3 * float(ref.sum())
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 129, 129, 257)   # contiguous, disjoint, covering


def test_single_process_allreduce_is_identity():
    from wavenet.train_step import allreduce_gradients, shard_streams
    g = torch.ones(8)
    assert allreduce_gradients(g) == 1.0 and float(g.sum()) == 8.0
    assert shard_streams(256, 3, 8) == (96, 128)


def test_workspace_of_the_benchmark_configs_fits_one_gpu():
    """Workspace sizes (host arithmetic only): BASELINE config 2 (default params, T = 100000) and config 5 (scaled net in
    16-bit storage, T = 65536) need a few GB of the 180 GB of a B200; forward-only needs less than training."""
    from wavenet import _lib
    lib = _lib.load()
    cfg2 = _lib.make_config([2 ** i for i in range(10)] * 5, 32, 32, 512, 256, None, None, True, False)
    cfg5 = _lib.make_config([2 ** i for i in range(10)] * 4, 128, 128, 512, 256, None, None, True, False)
    for cfg, t in ((cfg2, 100000), (cfg5, 65536)):
        train = lib.wn_train_workspace_bytes(C.byref(cfg), 1, t)
        fwd = lib.wn_forward_workspace_bytes(C.byref(cfg), 1, t)
        assert 1e9 < train < 8e9, train
        assert 0 < fwd < train
    # four windows per GPU scale the activation part
    assert lib.wn_train_workspace_bytes(C.byref(cfg2), 4, 100000) > 3 * lib.wn_train_workspace_bytes(C.byref(cfg2), 1, 100000)
