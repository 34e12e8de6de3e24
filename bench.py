#!/usr/bin/env python
"""Benchmark of the dilated-causal-convolution hot path (BASELINE.json: "training audio
samples/sec (default params); fast-gen samples/sec/stream").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full training step (mu-law encode -> network forward -> softmax-CE ->
backward -> gradient all-reduce (N>1) -> Adam) over one synthetic batch of the default
wavenet_params.json network at B=1 window of T=100000 samples per GPU (BASELINE config[1];
weak scaling: per-GPU work is fixed).  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, 'tensorflow-wavenet_b200'), os.path.join(ROOT, 'oracle')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

DEFAULT_PARAMS = json.load(open(os.path.join(ROOT, 'tensorflow-wavenet_b200', 'wavenet_params.json')))
METRIC = 'training audio samples/sec (default params)'
UNIT = 'samples/s'
T_WINDOW = 100000     # train.py:30 SAMPLE_SIZE
B_PER_GPU = 1         # train.py:22 BATCH_SIZE


def net_kwargs(batch):
    p = DEFAULT_PARAMS
    return dict(batch_size=batch, dilations=p['dilations'], filter_width=p['filter_width'],
                residual_channels=p['residual_channels'], dilation_channels=p['dilation_channels'],
                quantization_channels=p['quantization_channels'], skip_channels=p['skip_channels'],
                use_biases=p['use_biases'], scalar_input=p['scalar_input'],
                initial_filter_width=p['initial_filter_width'], residual_postproc=p['residual_postproc'])


def synthetic_audio(batch, time_steps, seed):
    """3-tone chord at 16 kHz + 0.05 N(0,1) noise, clipped to [-1, 1] (cfg 2)."""
    rng = np.random.default_rng(seed)
    t = np.arange(time_steps) / 16000.0
    chord = (np.sin(2 * np.pi * 155.56 * t) + np.sin(2 * np.pi * 196.0 * t) + np.sin(2 * np.pi * 233.08 * t)) / 3.0
    a = chord[None, :] + 0.05 * rng.standard_normal((batch, time_steps))
    return np.clip(a, -1.0, 1.0).astype(np.float32)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d['hbm_gbs'], tflops=d.get('bf16_tflops_sustained', d['bf16_tflops']), source='measured')
    return dict(hbm_gbs=6650.0, tflops=1590.0, source='fallback')


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40,
                 'hw_power_brake_slowdown': 0x80, 'sw_power_cap': 0x4}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle restatement of the reference's TF graph (op for op: dense one-hot,
# pad -> time_to_batch -> conv1d -> batch_to_time -> slice per conv, per-layer skip tensors,
# autograd backward, TF Adam) on the host cores.
# ------------------------------------------------------------------------------------------
def cpu_train_samples_per_sec(time_steps, steps, warmup, threads):
    import wavenet_oracle as O
    torch.set_num_threads(threads)
    kw = net_kwargs(1)
    net = O.OracleWaveNet(dtype=torch.float32, seed=0, faithful=True, **kw)
    opt = O.TFOptimizer('adam', 1e-3, 0.9)
    audio = synthetic_audio(1, time_steps, 0)
    params = net.state_dict()
    last = 'wavenet/dilated_stack/layer{}/dense'.format(len(kw['dilations']) - 1)

    def one():
        _, _, grads = net.loss_and_grads(audio)
        grads[last] = None
        grads[last + '_bias'] = None
        opt.apply(params, grads)
        net.load_state_dict(params)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return time_steps / dt, dt


def cpu_gen_samples_per_sec(n_samples, threads):
    import wavenet_oracle as O
    torch.set_num_threads(threads)
    net = O.OracleWaveNet(dtype=torch.float32, seed=0, **net_kwargs(1))
    rs = np.random.RandomState(0)
    net.init_ops()
    cur = 128
    t0 = time.perf_counter()
    for _ in range(n_samples):
        p = net.predict_proba_incremental(cur)
        p = O.scale_prediction(p, 1.0)
        cur = O.choice_from_uniform(p, rs.random_sample())
    return n_samples / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric/config.  TensorFlow 0.10 is not
    installable here, so this times the oracle port (kind "port") with all host threads; every step is
    a bounded sample (a shorter window of the same network), sized so the whole run ends in minutes."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # calibrate on a short window, then pick the window so that (steps+warmup) fits ~150 s
    _, dt_small = cpu_train_samples_per_sec(8192, 1, 1, threads)
    per_sample = dt_small / 8192
    budget = 150.0 / max(1, args.steps + args.warmup)
    t_ref = int(min(T_WINDOW, max(8192, budget / per_sample)))
    value, dt = cpu_train_samples_per_sec(t_ref, args.steps, args.warmup, threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'default wavenet_params.json training step, B=1, window of {} samples '
                               '(bounded sample of the T=100000 step)'.format(t_ref)},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': 'fwd+bwd+Adam on one {}-sample window per step, PyTorch-CPU restatement of the '
                                   'TF-0.10 graph (TensorFlow itself is not installable here)'.format(t_ref)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_cfg5(args):
    """BASELINE config 5: scaled net (res/dil 128, skip 512, 4 x dilations 1..512), 16-bit activation storage, 64k windows.
    Same contract as the default line (one JSON line, device-timed graph replays + an end-to-end leg); the roofline block
    is the tensor bound of the whole step: SURVEY 8(d) FLOPs per unit x units / step time against the measured bf16 peak."""
    import wavenet
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.distributed.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    B, T = 1, 65536
    R = D = 128
    S, Q = 512, 256
    dil = [2 ** i for i in range(10)] * 4
    L = len(dil)
    net = wavenet.WaveNetModel(batch_size=B, dilations=dil, filter_width=2, residual_channels=R, dilation_channels=D,
                               quantization_channels=Q, skip_channels=S, use_biases=True, seed=0)
    opt = wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9)
    step = wavenet.TrainStep(net, opt, B, T)
    host_audio = torch.as_tensor(synthetic_audio(B, T, rank)).pin_memory()
    step.audio.copy_(host_audio)
    W, K = max(3, args.warmup), max(1, args.steps)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(ms) / K
    value = world * B * T / (ms_per_step * 1e-3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        loss_host = float(step(host_audio))
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    if rank != 0:
        return
    peaks = measured_peaks()
    # forward + input-gradient + weight-gradient products: 3 x 2 x (2R.2D + D.R) per layer, 3 x 2 x (L.D.S + S.S + S.Q) post-processing
    flop_per_unit = 6.0 * (L * (2 * R * 2 * D + D * R) - D * R + L * D * S + S * S + S * Q)
    ach = flop_per_unit * B * T / (ms_per_step * 1e-3) / 1e12
    print(json.dumps({
        'metric': 'training audio samples/sec (scaled net, BASELINE config 5)', 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': K, 'warmup': W, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'fp16 activation / gradient storage and tensor-core operands, fp32 accumulation, parameters and optimizer',
        'data': 'synthetic',
        'config': {'workload': 'scaled net (L=%d, R=D=%d, S=%d, Q=%d, biases) training step (fwd+bwd+allreduce+Adam), B=%d x T=%d per GPU' % (L, R, S, Q, B, T),
                   'parallelism': 'dp%d' % world, 'l2_between_iterations': 'per-step working set (~4 GB) exceeds the 126 MB L2; no explicit flush'},
        'clocks': clocks,
        'e2e': {'value': world * B * T * K / float(e2e_s), 'unit': UNIT, 'h2d_bytes_per_step': B * T * 4, 'd2h_bytes_per_step': 4},
        'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': peaks['tflops'], 'unit': 'TFLOP/s', 'frac': ach / peaks['tflops'],
                     'traffic': None, 'kernel': 'whole step (fp16 tcgen05 GEMMs of the wide residual blocks + post-processing)',
                     'flop_per_unit': flop_per_unit, 'peak_source': peaks['source']},
        'final_loss': loss_host}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fastgen', action='store_true')
    ap.add_argument('--no-batch4', action='store_true')
    ap.add_argument('--batch', type=int, default=B_PER_GPU)
    ap.add_argument('--time', type=int, default=T_WINDOW)
    ap.add_argument('--config', default='default', choices=['default', 'cfg5'],
                    help='cfg5: BASELINE config 5, the scaled net (R = D = 128, 4 x dilations 1..512, 64k windows, 16-bit storage)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.config == 'cfg5':
        return run_cfg5(args)

    import wavenet
    from wavenet import _lib
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.distributed.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    B, T = args.batch, args.time

    trap = None
    if os.environ.get('WN_TRAP_INFO'):      # debug: identity of a bounded wait that timed out (wn_debug_trap_info)
        trap = torch.zeros(8, dtype=torch.int32).pin_memory()
        _lib.load().wn_debug_trap_info(C.c_void_p(trap.data_ptr()))
    net = wavenet.WaveNetModel(**net_kwargs(B), seed=0)      # same seed -> replicated weights
    opt = wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9)
    step = wavenet.TrainStep(net, opt, B, T)
    host_audio = torch.as_tensor(synthetic_audio(B, T, rank)).pin_memory()
    step.audio.copy_(host_audio)
    lib = _lib.load()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    try:
        for _ in range(W):
            step()
        barrier()
    except Exception:
        if trap is not None:
            print('trap info {bar_smem, parity, blockDim, gridDim.x, blockIdx, threadIdx, gridDim.y}:', trap.tolist(), flush=True)
        raise
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(ms) / K
    value = world * B * T / (ms_per_step * 1e-3)
    final_loss = float(step.loss)

    # ---------------- end to end: host (pinned) audio in, loss out, every step ----------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        loss_host = float(step(host_audio))     # H2D copy of the window + D2H read of the loss (syncs)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * B * T * K / float(e2e_s)

    # ---------------- per-kernel durations: CUDA events recorded after every launch INSIDE a captured graph
    # (an eager pass is host-launch-bound for the 10-30 us block kernels and would inflate them) ----------------
    n_tags = 23
    ms_tag = (C.c_float * n_tags)()
    n_tag = (C.c_int32 * n_tags)()
    prof_steps = 1
    prof_mode = 'graph'
    torch.cuda.synchronize()
    try:
        assert lib.wn_profile_begin() == 0
        gprof = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gprof):
            step._launch()
        for _ in range(3):
            gprof.replay()
        torch.cuda.synchronize()
        assert lib.wn_profile_end(ms_tag, n_tag, n_tags) == 0
    except Exception:
        prof_mode = 'eager'
        torch.cuda.synchronize()
        assert lib.wn_profile_begin() == 0
        step._launch()
        assert lib.wn_profile_end(ms_tag, n_tag, n_tags) == 0
    # cost of one event-record node between two kernel nodes: 64 back-to-back marks in a captured graph
    event_us = None
    if prof_mode == 'graph':
        try:
            cal_ms = (C.c_float * n_tags)()
            cal_n = (C.c_int32 * n_tags)()
            assert lib.wn_profile_begin() == 0
            gcal = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gcal):
                for _ in range(65):
                    lib.wn_profile_mark(0, _lib.stream_ptr())
            for _ in range(3):
                gcal.replay()
            torch.cuda.synchronize()
            assert lib.wn_profile_end(cal_ms, cal_n, n_tags) == 0
            event_us = 1e3 * cal_ms[0] / max(1, cal_n[0])
        except Exception:
            event_us = None
    kernels = {}
    buf = C.create_string_buffer(64)
    for i in range(n_tags):
        lib.wn_profile_tag_name(i, buf, 64)
        if n_tag[i]:
            kernels[buf.value.decode()] = {'launches_per_step': n_tag[i] // prof_steps,
                                           'ms_per_step': ms_tag[i] / prof_steps,
                                           'us_per_launch': 1e3 * ms_tag[i] / n_tag[i]}
    launches_per_step = sum(k['launches_per_step'] for k in kernels.values() if not k.get('stage')) + 1   # + optimizer kernel
    peaks = measured_peaks()
    M = B * T
    p = DEFAULT_PARAMS
    R, D, S, Q, L = (p['residual_channels'], p['dilation_channels'], p['skip_channels'],
                     p['quantization_channels'], len(p['dilations']))
    # algorithmic bytes / flops per launch
    algo = {
        # all forward layers (x in, x' and z out per layer; the last layer has no x'), divided by the launches that
        # cover them: one persistent kernel by default, L per-layer launches with WN_FWD_CHAIN=0
        'block_fwd': ('hbm', M * 4.0 * ((2 * R + D) * (L - 1) + (R + D)) / max(1, kernels.get('block_fwd', {}).get('launches_per_step', L))),
        # the WHOLE backward of the residual blocks (pre-activation + input gradients: one persistent kernel; weight
        # gradients: one launch) against SURVEY section 8(d)'s figure for the stage, (3R + 2D) * 4 = 640 B per unit per
        # layer -- every byte the two kernels move beyond that (dpre / dx round trips through L2 / HBM) counts against them
        'block_bwd': ('hbm', M * 4.0 * (3 * R + 2 * D) * L),
        'softmax_xent': ('hbm', M * 4.0 * 2 * Q),
        'gemm_skip_fwd': ('tensor', 2.0 * M * L * D * S),
        'gemm_skip_wgrad': ('tensor', 2.0 * M * L * D * S),
        'gemm_skip_dgrad': ('tensor', 2.0 * M * L * D * S),
        'gemm_post1_fwd': ('tensor', 2.0 * M * S * S),
        'gemm_post1_wgrad': ('tensor', 2.0 * M * S * S),
        'gemm_post1_dgrad': ('tensor', 2.0 * M * S * S),
        'gemm_post2_fwd': ('tensor', 2.0 * M * S * Q),
        'gemm_post2_wgrad': ('tensor', 2.0 * M * S * Q),
        'gemm_post2_dgrad': ('tensor', 2.0 * M * S * Q),
    }
    # stage entry: sum of the kernels that make up the block backward (tags block_bwd_pre [+ block_bwd_dx] + block_wgrad)
    bwd_parts = [k for k in ('block_bwd_pre', 'block_bwd_dx', 'block_wgrad') if k in kernels]
    if bwd_parts:
        ms = sum(kernels[k]['ms_per_step'] for k in bwd_parts)
        kernels['block_bwd'] = {'launches_per_step': 1, 'ms_per_step': ms, 'us_per_launch': 1e3 * ms,
                                'kernels': {k: kernels[k]['ms_per_step'] for k in bwd_parts}, 'stage': True}
    rooflines = {}
    for name, (bound, work) in algo.items():
        if name not in kernels:
            continue
        sec = kernels[name]['us_per_launch'] * 1e-6
        if bound == 'hbm':
            ach, peak, unit = work / sec / 1e9, peaks['hbm_gbs'], 'GB/s'
        else:
            ach, peak, unit = work / sec / 1e12, peaks['tflops'], 'TFLOP/s'
        net_sec = max(sec - (event_us or 0.0) * 1e-6, 1e-9)      # without the event node that follows the kernel
        ach_net = work / net_sec / (1e9 if bound == 'hbm' else 1e12)
        rooflines[name] = {'bound': bound, 'achieved': ach, 'peak': peak, 'unit': unit, 'frac': ach / peak,
                           'achieved_net_of_event_node': ach_net, 'frac_net_of_event_node': ach_net / peak,
                           'traffic': None, 'share_of_step': kernels[name]['ms_per_step'] /
                           sum(k['ms_per_step'] for k in kernels.values() if not k.get('stage'))}
    # measured DRAM traffic per launch of the dominant kernels (one `ncu --set full` capture, tools/ncu_summary.py traffic)
    tpath = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if os.path.exists(tpath):
        tr = json.load(open(tpath))
        for name in rooflines:
            if name in tr and args.batch == B_PER_GPU and args.time == T_WINDOW:
                rooflines[name]['traffic'] = tr[name]['dram_bytes_per_launch']
        if 'block_bwd' in rooflines and args.batch == B_PER_GPU and args.time == T_WINDOW and \
                'block_bwd_chain' in tr and 'block_wgrad' in tr:      # the stage = its two launches
            rooflines['block_bwd']['traffic'] = (tr['block_bwd_chain']['dram_bytes_per_launch'] +
                                                 tr['block_wgrad']['dram_bytes_per_launch'])
    dominant = max(rooflines, key=lambda n: kernels[n]['ms_per_step']) if rooflines else None
    roofline = dict(rooflines[dominant], kernel=dominant, peak_source=peaks['source'] +
                    (' (bf16 dense GEMM; the post-processing GEMMs run fp16 operands with fp32 accumulation, same nominal rate)'
                     if rooflines[dominant]['bound'] == 'tensor' else '')) if dominant else None

    # ---------------- the same step at B=4 windows per GPU (SURVEY section 8d "also report B=4/GPU") ----------------
    batch4 = None
    if not args.no_batch4 and args.batch == B_PER_GPU and args.time == T_WINDOW:
        try:
            net4 = wavenet.WaveNetModel(**net_kwargs(4), seed=0)
            step4 = wavenet.TrainStep(net4, wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9), 4, T)
            step4.audio.copy_(torch.as_tensor(synthetic_audio(4, T, 100 + rank)))
            for _ in range(3):
                step4()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(10):
                step4()
            f1.record()
            barrier()
            ms4 = torch.tensor([f0.elapsed_time(f1)], device=dev)
            if world > 1:
                torch.distributed.all_reduce(ms4, op=torch.distributed.ReduceOp.MAX)
            batch4 = {'value': world * 4 * T / (float(ms4) / 10 * 1e-3), 'unit': UNIT, 'ms_per_step': float(ms4) / 10,
                      'batch_per_gpu': 4}
            del step4, net4
            torch.cuda.empty_cache()
        except Exception as e:      # noqa: BLE001  (an optional extra must not take the headline down)
            batch4 = {'error': repr(e)[:200]}

    # ---------------- BASELINE config 5 (scaled net, 16-bit storage) next to the headline: `--config cfg5` gives the full line ----
    cfg5 = None
    if world == 1 and not args.no_batch4 and args.batch == B_PER_GPU and args.time == T_WINDOW:      # (N > 1: `--config cfg5` under torchrun)
        try:
            T5, R5, S5, Q5 = 65536, 128, 512, 256
            dil5 = [2 ** i for i in range(10)] * 4
            net5 = wavenet.WaveNetModel(batch_size=1, dilations=dil5, filter_width=2, residual_channels=R5, dilation_channels=R5,
                                        quantization_channels=Q5, skip_channels=S5, use_biases=True, seed=0)
            step5 = wavenet.TrainStep(net5, wavenet.optimizer_factory['adam'](learning_rate=1e-3, momentum=0.9), 1, T5)
            step5.audio.copy_(torch.as_tensor(synthetic_audio(1, T5, 200 + rank)))
            for _ in range(3):
                step5()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(10):
                step5()
            f1.record()
            barrier()
            ms5 = torch.tensor([f0.elapsed_time(f1)], device=dev)
            if world > 1:
                torch.distributed.all_reduce(ms5, op=torch.distributed.ReduceOp.MAX)
            L5 = len(dil5)
            flop5 = 6.0 * (L5 * (2 * R5 * 2 * R5 + R5 * R5) - R5 * R5 + L5 * R5 * S5 + S5 * S5 + S5 * Q5)
            tf5 = flop5 * T5 / (float(ms5) / 10 * 1e-3) / 1e12
            cfg5 = {'value': world * T5 / (float(ms5) / 10 * 1e-3), 'unit': UNIT, 'ms_per_step': float(ms5) / 10,
                    'workload': 'scaled net (L=40, R=D=128, S=512), fp16 activation storage, B=1 x T=65536 per GPU',
                    'roofline': {'bound': 'tensor', 'achieved': tf5, 'peak': measured_peaks()['tflops'], 'unit': 'TFLOP/s',
                                 'frac': tf5 / measured_peaks()['tflops']}}
            del step5, net5
            torch.cuda.empty_cache()
        except Exception as e:      # noqa: BLE001
            cfg5 = {'error': repr(e)[:200]}

    # ---------------- fast generation (second half of the BASELINE metric) ----------------
    fastgen = None
    if not args.no_fastgen:
        gnet = net
        gnet.batch_size = 1
        n1 = 16000          # BASELINE config 4: 16,000 samples per stream
        torch.cuda.synchronize()
        gnet.generate(64, [128], seed=0)                         # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gnet.generate(n1, [128], seed=0)
        torch.cuda.synchronize()
        b1 = n1 / (time.perf_counter() - t0)
        streams_total = 256
        from wavenet.train_step import shard_streams
        lo_s, hi_s = shard_streams(streams_total, rank, world)
        per_rank = hi_s - lo_s
        first = np.random.RandomState(rank).randint(0, 256, per_rank)
        n2 = 16000          # BASELINE config 4: 16,000 samples per stream, also for the sharded streams
        u = np.random.RandomState(100 + rank).random_sample((per_rank, n2))
        gnet.generate(16, first, uniforms=u[:, :16])
        barrier()
        t0 = time.perf_counter()
        gnet.generate(n2, first, uniforms=u)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
        fastgen = {'b1_samples_per_sec_per_stream': b1, 'b1_us_per_sample': 1e6 / b1,
                   'streams': streams_total, 'streams_per_gpu': per_rank,
                   'b256_samples_per_sec_per_stream': n2 / float(dt),
                   'b256_aggregate_samples_per_sec': streams_total * n2 / float(dt),
                   'kernel_256': ('pipelined layer-per-warp chain (2..32 streams per GPU)' if per_rank <= 32 else
                                  'throughput-mode kernel (one CTA per 1-4 streams)'),
                   'note': 'timed over {} (B=1, latency-mode kernel) / {} ({} streams per GPU) samples '
                           'incl. launch + H2D of the uniforms'.format(n1, n2, per_rank)}

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_train_samples_per_sec(T, 1, 0, threads)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': '1 training step (fwd+bwd+Adam) of the same B=1, T={} window, {:.1f} s; PyTorch-CPU '
                         'restatement of the TF-0.10 graph (TensorFlow not installable)'.format(T, dt)}
        if fastgen is not None:
            g = cpu_gen_samples_per_sec(300, threads)
            fastgen['cpu_b1_samples_per_sec_per_stream'] = g
            fastgen['b1_speedup_vs_cpu'] = fastgen['b1_samples_per_sec_per_stream'] / g

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 storage and accumulation; tensor-core operands fp16 split rows (residual blocks), fp16 / tf32 (GEMMs)', 'data': 'synthetic',
            'config': {'workload': 'default wavenet_params.json (L=50, R=D=32, S=512, Q=256, biases) training step '
                                   '(fwd+bwd+allreduce+Adam), B={} x T={} per GPU'.format(B, T),
                       'parallelism': 'dp{}'.format(world),
                       'l2_between_iterations': 'per-step working set (~4.6 GB of activations per GPU) exceeds the '
                                                '126 MB L2; no explicit flush'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * T * 4, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches_per_step * K,
            'launches_per_step': launches_per_step,
            'roofline': roofline,
            'roofline_all': rooflines,
            'kernels': kernels,
            'kernel_timing': prof_mode + ': CUDA events after every launch of one step (times include one event '
                             'node each; event_node_us = its cost measured with back-to-back marks)',
            'event_node_us': event_us,
            'cpu_baseline': cpu,
            'fastgen': fastgen,
            'batch4': batch4,
            'cfg5': cfg5,
            'loss': {'after_timed_steps': final_loss, 'e2e_last': loss_host},
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
