"""Host-side mirror of wavenet/model.py of the reference: the same constructor, attributes and
methods (`loss`, `predict_proba`, `predict_proba_incremental`, `variables`, `init_ops`,
`push_ops`, `batch_size`), executing on hand-written sm_100a kernels through the C ABI.

TensorFlow graph idioms map to eager calls:
  * `loss(...)` runs the fused forward+backward launch sequence and returns a 0-d CUDA tensor;
    the gradients of every variable are left in `flat_grads` (and `gradients()`), which is what
    `optimizer.minimize(loss)` consumes (reference: train.py:245-252).
  * `predict_proba_incremental(sample)` executes one generator step; `init_ops` / `push_ops`
    are lists of callables standing for the queue-initialisation / enqueue ops
    (model.py:490-491) so the "forward, then push" protocol of generate.py:195-226 and
    test/test_generation.py:61-69 stays expressible.
  * `generate(...)` is the batched, persistent-kernel form of the generate.py sample loop.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .ops import as_cuda, device_tables

_BIAS_AUTONAMES = {  # TF auto-names of the bias variables in this snapshot
    'filter_bias': 'Variable', 'gate_bias': 'Variable_1', 'dense_bias': 'Variable_2', 'slip_bias': 'Variable_3',
    'postprocess1_bias': 'Variable', 'postprocess2_bias': 'Variable_1'}


def _xavier(rng, shape):
    """tf.contrib.layers.xavier_initializer_conv2d (uniform), model.py:7-12."""
    recept = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    limit = math.sqrt(6.0 / ((shape[-2] + shape[-1]) * recept))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


class WaveNetModel(object):
    '''Implements the WaveNet network for generative audio (reference model.py:31-685).

    Usage (with the architecture as in the DeepMind paper):
        dilations = [2**i for i in range(N)] * M
        net = WaveNetModel(batch_size, dilations, filter_width=2, residual_channels=32,
                           dilation_channels=32, skip_channels=512)
        loss = net.loss(input_batch)
    '''

    def __init__(self, batch_size, dilations, filter_width, residual_channels, dilation_channels,
                 skip_channels, quantization_channels=2 ** 8, use_biases=False, scalar_input=False,
                 initial_filter_width=32, histograms=False, global_condition_channels=None,
                 global_condition_cardinality=None, residual_postproc=False, seed=None):
        self.batch_size = batch_size
        self.dilations = list(dilations)
        self.filter_width = filter_width
        self.residual_channels = residual_channels
        self.dilation_channels = dilation_channels
        self.quantization_channels = quantization_channels
        self.use_biases = use_biases
        self.skip_channels = skip_channels
        self.scalar_input = scalar_input
        self.initial_filter_width = initial_filter_width
        self.histograms = histograms
        self.global_condition_channels = global_condition_channels
        self.global_condition_cardinality = global_condition_cardinality
        self.residual_postproc = residual_postproc

        _lib.require_cuda()
        self._lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._native = (filter_width == 2)      # (filter_width > 2 exists only as the ops.causal_conv op)
        self._workspaces = {}
        self._gen = None
        self.init_ops = [self._init_generator]
        self.push_ops = [self._push_generator]
        self.variables = self._create_variables(seed)

    # ------------------------------------------------------------------ variables
    def _create_variables(self, seed):
        '''Creates all variables (model.py:118-225) as views into one flat fp32 CUDA buffer
        whose group layout is dictated by the C ABI (wn_param_layout).'''
        rng = np.random.default_rng(seed)
        L, R, D = len(self.dilations), self.residual_channels, self.dilation_channels
        S, Q, G = self.skip_channels, self.quantization_channels, self.global_condition_channels
        card = self.global_condition_cardinality
        if not self._native:
            # filter_width > 2 has no block kernels: the object is constructible like the reference, loss/predict raise.
            self._cfg = self._layout = None
            self.flat_params = self.flat_grads = None
            return None
        self._cfg = _lib.make_config(self.dilations, R, D, S, Q, G, card, self.use_biases, self.residual_postproc,
                                     self.scalar_input, self.initial_filter_width)
        self._layout = lo = _lib.param_layout(self._cfg)
        self.flat_params = torch.zeros(lo.total, dtype=torch.float32, device=self.device)
        self.flat_grads = torch.zeros(lo.total, dtype=torch.float32, device=self.device)

        def view(buf, off, shape):
            n = int(np.prod(shape))
            return buf[off:off + n].view(*shape)

        self._groups = {}   # group name -> (offset, per-layer shape or full shape, layered?)
        def group(name, off, shape, layered):
            if off >= 0:
                self._groups[name] = (off, tuple(shape), layered)

        # model.py:141-154: [2, Q, R] on the one-hot encoding, [initial_filter_width, 1, R] on the scalar waveform
        group('causal', lo.causal, (self.initial_filter_width, 1, R) if self.scalar_input else (2, Q, R), False)
        group('filter', lo.filter, (2, R, D), True)
        group('gate', lo.gate, (2, R, D), True)
        group('dense', lo.dense, (1, D, R), True)
        group('skip', lo.skip, (1, D, S), True)
        if G:
            group('gc_filter', lo.gc_filter, (1, G, D), True)
            group('gc_gate', lo.gc_gate, (1, G, D), True)
        group('filter_bias', lo.filter_bias, (D,), True)
        group('gate_bias', lo.gate_bias, (D,), True)
        group('dense_bias', lo.dense_bias, (R,), True)
        group('skip_bias', lo.skip_bias, (S,), True)
        group('post1', lo.post1, (1, S, S), False)
        group('post2', lo.post2, (1, S, Q), False)
        group('post1_bias', lo.post1_bias, (S,), False)
        group('post2_bias', lo.post2_bias, (Q,), False)
        if G and card:
            group('gc_embedding', lo.gc_embedding, (card, G), False)

        def v(buf, name, layer=None):
            off, shape, layered = self._groups[name]
            n = int(np.prod(shape))
            if layered:
                off += layer * n
            return view(buf, off, shape)

        self._view = v
        var, grad = dict(), dict()
        for tree, buf in ((var, self.flat_params), (grad, self.flat_grads)):
            if 'gc_embedding' in self._groups:
                tree['embeddings'] = {'gc_embedding': v(buf, 'gc_embedding')}
            tree['causal_layer'] = {'filter': v(buf, 'causal')}
            tree['dilated_stack'] = []
            for i in range(L):
                cur = {'filter': v(buf, 'filter', i), 'gate': v(buf, 'gate', i),
                       'dense': v(buf, 'dense', i), 'skip': v(buf, 'skip', i)}
                if G:
                    cur['gc_gateweights'] = v(buf, 'gc_gate', i)
                    cur['gc_filtweights'] = v(buf, 'gc_filter', i)
                if self.use_biases:
                    cur['filter_bias'] = v(buf, 'filter_bias', i)
                    cur['gate_bias'] = v(buf, 'gate_bias', i)
                    cur['dense_bias'] = v(buf, 'dense_bias', i)
                    cur['skip_bias'] = v(buf, 'skip_bias', i)
                tree['dilated_stack'].append(cur)
            post = {'postprocess1': v(buf, 'post1'), 'postprocess2': v(buf, 'post2')}
            if self.use_biases:
                post['postprocess1_bias'] = v(buf, 'post1_bias')
                post['postprocess2_bias'] = v(buf, 'post2_bias')
            tree['postprocessing'] = post
        self._grad_tree = grad

        # initial values in the reference's creation order (model.py:126-222); biases are zero
        with torch.no_grad():
            def put(t):
                t.copy_(torch.as_tensor(_xavier(rng, tuple(t.shape)), device=self.device))
            if 'embeddings' in var:
                if card == G:
                    var['embeddings']['gc_embedding'].copy_(torch.eye(card, device=self.device))
                else:
                    put(var['embeddings']['gc_embedding'])
            put(var['causal_layer']['filter'])
            for cur in var['dilated_stack']:
                for key in ('filter', 'gate', 'dense', 'skip', 'gc_gateweights', 'gc_filtweights'):
                    if key in cur:
                        put(cur[key])
            put(var['postprocessing']['postprocess1'])
            put(var['postprocessing']['postprocess2'])
        return var

    def _named(self, tree):
        out = {}
        if 'embeddings' in tree:
            out['wavenet/embeddings/gc_embedding'] = tree['embeddings']['gc_embedding']
        out['wavenet/causal_layer/filter'] = tree['causal_layer']['filter']
        names = (('filter', 'filter'), ('gate', 'gate'), ('dense', 'dense'), ('skip', 'skip'),
                 ('gc_gateweights', 'gc_gate'), ('gc_filtweights', 'gc_filter'), ('filter_bias', 'filter_bias'),
                 ('gate_bias', 'gate_bias'), ('dense_bias', 'dense_bias'), ('skip_bias', 'slip_bias'))
        for i, cur in enumerate(tree['dilated_stack']):
            for key, nm in names:
                if key in cur:
                    out['wavenet/dilated_stack/layer{}/{}'.format(i, nm)] = cur[key]
        for key, t in tree['postprocessing'].items():
            out['wavenet/postprocessing/' + key] = t
        return out

    def state_dict(self):
        """Checkpoint view: reference variable names (intended bias names) -> numpy."""
        self._require_native()
        return {k: t.detach().cpu().numpy().copy() for k, t in self._named(self.variables).items()}

    def load_state_dict(self, sd):
        """Accepts the intended names and the TF auto-names ('.../Variable_2') of the biases."""
        self._require_native()
        with torch.no_grad():
            for name, t in self._named(self.variables).items():
                src = sd.get(name)
                if src is None:
                    scope, leaf = name.rsplit('/', 1)
                    if leaf in _BIAS_AUTONAMES:
                        src = sd.get(scope + '/' + _BIAS_AUTONAMES[leaf])
                if src is None:
                    raise KeyError('missing variable {} in state dict'.format(name))
                a = np.asarray(src, dtype=np.float32)
                if tuple(a.shape) != tuple(t.shape):
                    raise ValueError('shape mismatch for {}: {} vs {}'.format(name, a.shape, tuple(t.shape)))
                t.copy_(torch.as_tensor(a, device=self.device))

    def gradients(self):
        """name -> numpy gradient of the last loss() call (d loss / d variable)."""
        return {k: t.detach().cpu().numpy().copy() for k, t in self._named(self._grad_tree).items()}

    # ------------------------------------------------------------------ helpers
    def _require_native(self):
        if not self._native:
            raise NotImplementedError('filter_width > 2 has no sm_100a block kernels in this build (the causal_conv op '
                                      'takes any width); SURVEY section 8, row f4')

    def _workspace(self, kind, batch, time):
        key = (kind, batch, time)
        ws = self._workspaces.get(key)
        if ws is None:
            fn = self._lib.wn_train_workspace_bytes if kind == 'train' else self._lib.wn_forward_workspace_bytes
            nbytes = fn(C.byref(self._cfg), batch, time)
            if nbytes < 0:
                raise ValueError('invalid batch/time {}x{}'.format(batch, time))
            for k in [k for k in self._workspaces if k[0] == kind]:
                del self._workspaces[k]
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._workspaces[key] = ws
        return ws

    def _gc_ids(self, global_condition, count):
        if self.global_condition_channels is None:
            return None
        if global_condition is None:
            return None
        if self.global_condition_cardinality is None:
            raise NotImplementedError('conditioning on an already-embedded dense vector is not supported '
                                      '(the reference branch, model.py:541-555, is itself broken)')
        ids = as_cuda(global_condition, torch.int32).reshape(-1)
        if ids.numel() == 1 and count > 1:
            ids = ids.expand(count).contiguous()
        if ids.numel() != count:
            raise ValueError('Shape of global_condition {} does not match batch size {}.'.format(
                tuple(ids.shape), count))
        # tf.nn.embedding_lookup raises on the CPU for an id outside the table (model.py:539-540); the kernels
        # themselves treat such an id as a zero embedding and never touch memory outside the table
        if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= self.global_condition_cardinality):
            raise ValueError('global_condition id outside [0, {})'.format(self.global_condition_cardinality))
        return ids

    # ------------------------------------------------------------------ training
    def loss(self, input_batch, global_condition_batch=None, l2_regularization_strength=None, name='wavenet'):
        '''Creates a WaveNet network and returns the autoencoding loss (model.py:628-685).

        input_batch: float audio [B,T,1] / [B,T] / [T] (B == batch_size).  Side effect: the gradient
        of the returned loss w.r.t. every variable is left in `flat_grads`.'''
        self._require_native()
        audio = as_cuda(input_batch, torch.float32).reshape(self.batch_size, -1)
        B, T = audio.shape
        gc = self._gc_ids(global_condition_batch, B)
        if self.global_condition_channels is not None and gc is None:
            raise ValueError('global_condition_batch is required for a globally conditioned model')
        ws = self._workspace('train', B, T)
        thr, _ = device_tables(self.quantization_channels)
        out = torch.empty((), dtype=torch.float32, device=self.device)
        rc = self._lib.wn_loss_grad(C.byref(self._cfg), _lib.ptr(self.flat_params), _lib.ptr(self.flat_grads),
                                    _lib.ptr(ws), ws.numel(), _lib.ptr(audio), _lib.ptr(gc), _lib.ptr(thr), B, T,
                                    _lib.ptr(out), _lib.stream_ptr())
        _lib.check(rc, 'wn_loss_grad')
        l2 = 0.0
        if l2_regularization_strength is not None:
            # model.py:670-680: sum of tf.nn.l2_loss over ALL trainables in this snapshot (App. A10);
            # alignment gaps of the flat buffer are zero.  The gradient term is added by the optimizer.
            l2 = float(l2_regularization_strength)
            _lib.check(self._lib.wn_add_l2(_lib.ptr(out), _lib.ptr(self.flat_params), self.flat_params.numel(), l2,
                                           _lib.stream_ptr()), 'wn_add_l2')
        out._wavenet_model = self
        out._wavenet_l2 = l2
        return out

    # ------------------------------------------------------------------ summaries (model.py:314-325,668,682-683)
    def summaries(self, loss=None, bins=30):
        """What the reference hands to TensorBoard, as plain Python data: the scalar 'loss' (model.py:668) / 'total_loss'
        (:682-683, when the loss tensor carries an L2 term) and, with `histograms=True`, one histogram per layer weight /
        bias (:314-325: 'layer{i}_filter', '_gate', '_dense', '_skip', '_biases_filter', ...) as (counts, bin edges)."""
        out = {}
        if loss is not None:
            l2 = getattr(loss, '_wavenet_l2', 0.0)
            total = float(loss)
            if l2:
                out['total_loss'] = total
                out['loss'] = total - l2 * 0.5 * float((self.flat_params.double() ** 2).sum())
            else:
                out['loss'] = total
        if self.histograms:
            names = (('filter', 'filter'), ('gate', 'gate'), ('dense', 'dense'), ('skip', 'skip'),
                     ('filter_bias', 'biases_filter'), ('gate_bias', 'biases_gate'), ('dense_bias', 'biases_dense'),
                     ('skip_bias', 'biases_skip'))
            for i, cur in enumerate(self.variables['dilated_stack']):
                for key, tag in names:
                    if key in cur:
                        counts, edges = np.histogram(cur[key].detach().cpu().numpy().ravel(), bins=bins)
                        out['layer{}_{}'.format(i, tag)] = (counts, edges)
        return out

    # ------------------------------------------------------------------ naive prediction
    def _net_input(self, waveform):
        """Encoded ids [B, T] -> what the network reads: the ids themselves, or (scalar_input) their mu-law decoded
        float values (model.py:570-576)."""
        ids = as_cuda(waveform, torch.int32).reshape(self.batch_size, -1)
        if self.scalar_input:
            from .ops import mu_law_decode
            return mu_law_decode(ids, self.quantization_channels)
        return ids

    def _logits(self, ids, gc):
        B, T = ids.shape
        ws = self._workspace('fwd', B, T)
        logits = torch.empty((B * T, self.quantization_channels), dtype=torch.float32, device=self.device)
        rc = self._lib.wn_forward_logits(C.byref(self._cfg), _lib.ptr(self.flat_params), _lib.ptr(ws), ws.numel(),
                                         _lib.ptr(ids), _lib.ptr(gc), B, T, _lib.ptr(logits), _lib.stream_ptr())
        _lib.check(rc, 'wn_forward_logits')
        return logits

    def logits(self, waveform, global_condition=None):
        """Raw network output [B, T, Q] for encoded input ids (model.py:389-442)."""
        self._require_native()
        ids = self._net_input(waveform)
        gc = self._gc_ids(global_condition, ids.shape[0])
        return self._logits(ids, gc).view(ids.shape[0], ids.shape[1], -1)

    def predict_proba(self, waveform, global_condition=None, name='wavenet'):
        '''Computes the probability distribution of the next sample based on all samples in the
        input waveform (model.py:564-590): float64 softmax of the last row, returned as float32.'''
        self._require_native()
        ids = self._net_input(waveform)
        B, T = ids.shape
        gc = self._gc_ids(global_condition, B)
        ws = self._workspace('fwd', B, T)
        proba = torch.empty((self.quantization_channels,), dtype=torch.float32, device=self.device)
        rc = self._lib.wn_predict_last(C.byref(self._cfg), _lib.ptr(self.flat_params), _lib.ptr(ws), ws.numel(),
                                       _lib.ptr(ids), _lib.ptr(gc), B, T, _lib.ptr(proba), _lib.stream_ptr())
        _lib.check(rc, 'wn_predict_last')
        return proba

    # ------------------------------------------------------------------ fast generation
    def _gen_state(self, streams):
        if self.scalar_input:
            raise NotImplementedError("Scalar input is not supported by fast generation.")
        if self.dilation_channels != self.residual_channels:
            raise NotImplementedError('fast generation needs dilation_channels == residual_channels in this '
                                      'build (training / predict_proba take any widths)')
        if self._gen is None or self._gen['streams'] != streams:
            nbytes = self._lib.wn_gen_state_bytes(C.byref(self._cfg), streams)
            if nbytes < 0:
                raise ValueError('invalid stream count {}'.format(streams))
            self._gen = dict(streams=streams, state=torch.empty(nbytes, dtype=torch.uint8, device=self.device),
                             ready=False)
        return self._gen

    def _init_generator(self, streams=None):
        """init_ops (model.py:457-463,477-484): every delay line filled with zeros."""
        self._require_native()
        g = self._gen_state(streams or (self._gen['streams'] if self._gen else self.batch_size))
        _lib.check(self._lib.wn_gen_reset(C.byref(self._cfg), _lib.ptr(g['state']), g['streams'],
                                          _lib.stream_ptr()), 'wn_gen_reset')
        g['ready'] = True

    def _push_generator(self):
        """push_ops (model.py:461,482): enqueue what the last predict_proba_incremental computed."""
        g = self._gen
        if g is None or not g['ready']:
            raise RuntimeError('push_ops before predict_proba_incremental')
        _lib.check(self._lib.wn_gen_commit(C.byref(self._cfg), _lib.ptr(g['state']), g['streams'],
                                           _lib.stream_ptr()), 'wn_gen_commit')

    def predict_proba_incremental(self, waveform, global_condition=None, name='wavenet'):
        '''Computes the probability distribution of the next sample incrementally, based on a
        single sample and all previously passed samples (model.py:592-626).  Returns [Q] (the last
        stream's distribution, like the reference's last-row slice); does NOT advance the delay
        lines -- run `push_ops` for that.'''
        if self.filter_width > 2:
            raise NotImplementedError("Incremental generation does not support filter_width > 2.")
        if self.scalar_input:
            raise NotImplementedError("Scalar input is not supported by fast generation.")
        ids = as_cuda(waveform, torch.int32).reshape(-1)
        streams = ids.numel()
        g = self._gen_state(streams)
        if not g['ready']:
            self._init_generator(streams)
        gc = self._gc_ids(global_condition, streams)
        proba = torch.empty((streams, self.quantization_channels), dtype=torch.float32, device=self.device)
        rc = self._lib.wn_gen_run(C.byref(self._cfg), _lib.ptr(self.flat_params), _lib.ptr(g['state']), streams,
                                  _lib.ptr(ids), None, _lib.ptr(gc), None, 1, 1.0, 0, None, _lib.ptr(proba),
                                  _lib.stream_ptr())
        _lib.check(rc, 'wn_gen_run')
        return proba[-1]

    def generate(self, n_samples, first_samples, global_condition=None, temperature=1.0, uniforms=None,
                 seed=None, reset=True, return_proba=False):
        """Batched fast generation: the generate.py:213-241 loop for `len(first_samples)` independent
        streams inside one persistent kernel.  `uniforms` [streams, n_samples] float64 are the draws
        np.random.random_sample() would make (default: MT19937 seeded with `seed` + stream index).
        Returns int32 [streams, n_samples] (and the last distribution when return_proba)."""
        self._require_native()
        ids = as_cuda(first_samples, torch.int32).reshape(-1)
        streams = ids.numel()
        g = self._gen_state(streams)
        if reset or not g['ready']:
            self._init_generator(streams)
        gc = self._gc_ids(global_condition, streams)
        if uniforms is None:
            # seed=None: like generate.py:237-238, draws come from the unseeded global np.random stream
            base = int(np.random.randint(0, 2 ** 31 - 1 - streams)) if seed is None else int(seed)
            uniforms = np.stack([np.random.RandomState(base + s).random_sample(n_samples) for s in range(streams)])
        u = as_cuda(uniforms, torch.float64).reshape(streams, n_samples)
        out = torch.empty((streams, n_samples), dtype=torch.int32, device=self.device)
        proba = (torch.empty((streams, self.quantization_channels), dtype=torch.float32, device=self.device)
                 if return_proba else None)
        rc = self._lib.wn_gen_run(C.byref(self._cfg), _lib.ptr(self.flat_params), _lib.ptr(g['state']), streams,
                                  _lib.ptr(ids), None, _lib.ptr(gc), _lib.ptr(u), int(n_samples), float(temperature),
                                  1, _lib.ptr(out), _lib.ptr(proba), _lib.stream_ptr())
        _lib.check(rc, 'wn_gen_run')
        return (out, proba) if return_proba else out

    def prime(self, waveforms, global_condition=None, reset=True):
        """Feed known samples through the generator (generate.py:195-210 priming loop) in one launch:
        waveforms int [streams, n]; returns the distribution after the last one [streams, Q]."""
        self._require_native()
        w = as_cuda(waveforms, torch.int32)
        if w.dim() == 1:
            w = w.reshape(1, -1)
        streams, n = w.shape
        g = self._gen_state(streams)
        if reset or not g['ready']:
            self._init_generator(streams)
        gc = self._gc_ids(global_condition, streams)
        proba = torch.empty((streams, self.quantization_channels), dtype=torch.float32, device=self.device)
        rc = self._lib.wn_gen_run(C.byref(self._cfg), _lib.ptr(self.flat_params), _lib.ptr(g['state']), streams,
                                  None, _lib.ptr(w), _lib.ptr(gc), None, n, 1.0, 1, None, _lib.ptr(proba),
                                  _lib.stream_ptr())
        _lib.check(rc, 'wn_gen_run')
        return proba
