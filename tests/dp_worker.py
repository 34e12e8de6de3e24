"""Worker of tests/test_gpu_parity_full.py::test_data_parallel_two_ranks_equals_one_rank_batch_two (run under torchrun,
one rank per GPU, NCCL).  Rank r trains on window r; checked: the all-reduced mean gradient equals the gradient one
GPU computes for the batch of both windows, and the replicated weights stay bit-identical over 5 Adam steps."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-wavenet_b200'))


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl')
    import wavenet
    from wavenet.train_step import allreduce_gradients
    kw = dict(dilations=[2 ** i for i in range(10)] * 2, filter_width=2, residual_channels=32, dilation_channels=32,
              quantization_channels=256, skip_channels=512, use_biases=True)
    T = 5000
    rng = np.random.default_rng(0)
    audio = np.clip(0.4 * np.sin(np.arange(T) * 0.07)[None] + 0.2 * rng.standard_normal((world, T)), -1, 1).astype(np.float32)
    net = wavenet.WaveNetModel(batch_size=1, seed=5, **kw)
    # 1) gradients: mean over ranks of the per-window gradients == gradient of the global batch on one GPU
    net.loss(audio[rank])
    scale = allreduce_gradients(net.flat_grads)
    g_dp = (net.flat_grads * scale).cpu().numpy()
    big = wavenet.WaveNetModel(batch_size=world, seed=5, **kw)
    big.loss(audio)
    g_one = big.flat_grads.cpu().numpy()
    err = float(np.abs(g_dp - g_one).max() / np.abs(g_one).max())
    assert err < 1e-5, err
    # 2) replicas stay bit-identical: 5 graph-replayed steps, the bucketed all-reduce captured INSIDE the graph (the tail
    #    bucket reduced while the backward chain runs) -- and they equal the plain schedule (one all-reduce after the replay)
    #    The bucketed schedule runs two collectives of ONE communicator on two streams of the same graph, which NCCL does not
    #    guarantee to be deadlock-free next to other in-flight collectives: it is exercised only on request
    #    (WN_TEST_DP_OVERLAP=1); by default both steps use the plain schedule and must agree bit for bit.
    opt = wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9)
    bucketed = os.environ.get('WN_TEST_DP_OVERLAP', '0') == '1'
    if bucketed:
        os.environ['WN_DP_OVERLAP'] = '1'
    step = wavenet.TrainStep(net, opt, 1, T)
    assert step.overlap == bucketed
    os.environ.pop('WN_DP_OVERLAP', None)
    net_plain = wavenet.WaveNetModel(batch_size=1, seed=5, **kw)
    step_plain = wavenet.TrainStep(net_plain, wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9), 1, T)
    assert not step_plain.overlap      # the default: one all-reduce of the whole buffer after the graph replay
    for i in range(5):
        step(np.roll(audio[rank], 17 * i))
        step_plain(np.roll(audio[rank], 17 * i))
    torch.cuda.synchronize()
    mine = net.flat_params.clone()
    for params in (mine, net_plain.flat_params.clone()):      # both schedules keep the replicas bit-identical
        gathered = [torch.empty_like(params) for _ in range(world)]
        dist.all_gather(gathered, params)
        for other in gathered:
            assert torch.equal(other, params), 'replicas diverged'
    moved = float((mine - big.flat_params).abs().max())
    assert moved > 0
    sched = float((mine - net_plain.flat_params).abs().max())
    assert sched <= 2e-4, sched      # (two runs order their fp32 atomics differently, and Adam normalises tiny gradients)
    dist.barrier()
    if rank == 0:
        print('DP_OK grad max-norm rel err {:.2e}; params moved by {:.2e}; overlapped vs plain schedule {:.1e}'.format(err, moved, sched))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
