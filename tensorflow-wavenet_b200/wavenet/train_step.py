"""One data-parallel training step (reference: the `sess.run([loss, optim])` of train.py:287-306).

Every rank runs the full model on its own [B_local, T] windows (the reference is single-device;
data parallelism is this build's addition).  The only exchange is a sum
all-reduce of the flat fp32 gradient buffer over NCCL, followed by the same optimizer update on
every rank (weights stay replicated).

The all-reduce is split into two buckets and lives INSIDE the captured CUDA graph of the step:
  * tail bucket [layout.skip, layout.total): skip / postprocess1 / postprocess2 weights and biases, 80 % of the bytes.
    Their gradients are final when the post-processing gradient GEMMs are done -- more than a millisecond before the
    end of the step; the library records an event there (wn_set_grad_ready_event) and the bucket is reduced on a
    communication stream while the residual-block backward still runs;
  * head bucket [0, layout.skip): everything the residual-block backward produces, reduced at the end.
Measured on 2 x B200 (gpurun_out/r2_dp_bench_*.log): 2.805 ms / step with the bucketed in-graph all-reduce against
2.767 ms with ONE all-reduce of the whole buffer after the graph replay -- the persistent backward kernels occupy every
SM with two CTAs, so the NCCL kernel of the tail bucket does not get an SM before they are done, and the second
collective only adds launch latency.  The bucketed form is therefore OFF by default (WN_DP_OVERLAP=1 turns it on).  It runs
two collectives of one communicator on two streams of one graph, which NCCL does not guarantee to be deadlock-free next to
other in-flight collectives: the 2-rank test exercises it only with WN_TEST_DP_OVERLAP=1.
"""
import ctypes as C
import os

import torch

from . import _lib
from .ops import as_cuda, device_tables


def allreduce_gradients(flat_grads, group=None):
    """Sum the flat gradient buffer over all ranks (NCCL on GPUs, any backend in tests) and return the factor
    the optimizer must scale it by so that the update uses the MEAN over ranks -- with equal B*T per rank that
    is the reference's reduce_mean over the global batch (model.py:666)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return 1.0
    world = torch.distributed.get_world_size(group)
    if world == 1:
        return 1.0
    torch.distributed.all_reduce(flat_grads, op=torch.distributed.ReduceOp.SUM, group=group)
    return 1.0 / world


def shard_streams(n_streams, rank, world):
    """Generation streams are independent: rank r generates streams [lo, hi) and no collective is needed
    (SURVEY section 8e).  Remainders go to the lowest ranks."""
    base, rem = divmod(int(n_streams), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class TrainStep(object):
    def __init__(self, net, optimizer, batch, time, l2_regularization_strength=None, process_group=None,
                 use_cuda_graph=True):
        net._require_native()
        self.net, self.opt = net, optimizer
        self.batch, self.time = int(batch), int(time)
        self.l2 = 0.0 if l2_regularization_strength is None else float(l2_regularization_strength)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        dev = net.device
        self.audio = torch.zeros((self.batch, self.time), dtype=torch.float32, device=dev)   # static inputs
        self.gc = (torch.zeros((self.batch,), dtype=torch.int32, device=dev)
                   if net.global_condition_channels is not None else None)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self._ws = net._workspace('train', self.batch, self.time)
        self._thr, _ = device_tables(net.quantization_channels)
        self._graph = None
        self.kernel_launches = None
        # bucketed all-reduce inside the step (world > 1, NCCL): see the module docstring
        self.overlap = (self.world > 1 and os.environ.get('WN_DP_OVERLAP', '0') == '1' and
                        torch.distributed.get_backend(process_group) == 'nccl')
        self._reduced_in_step = False
        if self.overlap:
            self._comm = torch.cuda.Stream(device=dev)
            self._ev_tail = torch.cuda.Event()
            self._ev_comm = torch.cuda.Event()
            self._ev_tail.record()                      # (creates the CUDA event: its handle goes to the library)
            self._tail_off = int(net._layout.skip)
            torch.distributed.all_reduce(torch.zeros(8, device=dev), group=process_group)   # communicator up before any capture
        if use_cuda_graph:
            self._capture()

    def _launch(self):
        n = self.net
        if self.overlap:
            _lib.check(n._lib.wn_set_grad_ready_event(C.c_void_p(self._ev_tail.cuda_event)), 'wn_set_grad_ready_event')
        try:
            rc = n._lib.wn_loss_grad(C.byref(n._cfg), _lib.ptr(n.flat_params), _lib.ptr(n.flat_grads),
                                     _lib.ptr(self._ws), self._ws.numel(), _lib.ptr(self.audio), _lib.ptr(self.gc),
                                     _lib.ptr(self._thr), self.batch, self.time, _lib.ptr(self.loss), _lib.stream_ptr())
        finally:
            if self.overlap:
                n._lib.wn_set_grad_ready_event(None)
        _lib.check(rc, 'wn_loss_grad')
        if self.overlap:
            cur = torch.cuda.current_stream()
            g = n.flat_grads
            self._comm.wait_event(self._ev_tail)        # recorded by the library where the tail bucket's gradients are final
            with torch.cuda.stream(self._comm):
                torch.distributed.all_reduce(g[self._tail_off:], op=torch.distributed.ReduceOp.SUM, group=self.pg)
                self._ev_comm.record(self._comm)
            torch.distributed.all_reduce(g[:self._tail_off], op=torch.distributed.ReduceOp.SUM, group=self.pg)
            cur.wait_event(self._ev_comm)
            self._reduced_in_step = True

    def _capture(self):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._launch()          # warm-up: function attributes, lazy module load
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch()
        self._graph = g

    def __call__(self, audio=None, gc_ids=None):
        """Runs one step on `audio` [B,T] (host or device; copied into the static input buffer)
        and returns the (device) loss of this rank."""
        if audio is not None:
            src = audio if isinstance(audio, torch.Tensor) else torch.as_tensor(audio)
            self.audio.copy_(src.reshape(self.batch, self.time), non_blocking=True)
        if gc_ids is not None and self.gc is not None:
            src = gc_ids if isinstance(gc_ids, torch.Tensor) else torch.as_tensor(gc_ids)
            self.gc.copy_(src.reshape(-1).to(torch.int32), non_blocking=True)
        if self._graph is not None:
            self._graph.replay()
        else:
            self._launch()
        if self._reduced_in_step:
            scale = 1.0 / self.world
        else:
            scale = allreduce_gradients(self.net.flat_grads, self.pg) if self.world > 1 else 1.0
        self.opt.apply(self.net.flat_params, self.net.flat_grads, l2=self.l2, grad_scale=scale)
        return self.loss
