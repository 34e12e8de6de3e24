"""CPU oracle for the tensorflow-wavenet hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU (NumPy for the integer / companding work, PyTorch-CPU
tensors for the float network so that autograd can give gradient oracles), the
arithmetic that jyegerlehner/tensorflow-wavenet asks TensorFlow 0.10 to perform on
its hot path.  Nothing under ``tensorflow-wavenet_b200/`` may import it: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do, and there only as the checker / the timed CPU stand-in.

Pinning status
--------------
* mu-law encode/decode and ``causal_conv`` are PINNED by the reference's own golden
  vectors (``test/test_mu_law.py:113-124`` known-answer vector, ``:126-178`` seeded
  equalities against its float32 numpy formulas, ``test/test_causal_conv.py:11-58``)
  -- see ``tests/test_oracle_golden.py``.
* Network logits / loss / gradients: **parity unpinned**.  The arithmetic lives in
  TensorFlow 0.10.0 (pinned in ``.travis.yml:7,9`` / ``ci/install.sh:13-17``; not in
  ``requirements.txt``), which is neither vendored under /root/reference nor
  installable here, and the reference ships no golden logits or losses.  The only
  reference-side pins are self-consistency ones (naive == incremental,
  ``test/test_generation.py:50-72``; loss thresholds ``test/test_model.py:275-282``),
  which the oracle reproduces (``tests/test_oracle_network.py``).

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# mu-law companding                                   wavenet/ops.py:65-85
# --------------------------------------------------------------------------- #


def mu_law_encode(audio, quantization_channels):
    """ops.py:65-73.  Every intermediate is float32 (TF graph of float32 tensors with
    Python scalars folded to float32 constants); the final cast truncates toward 0."""
    f32 = np.float32
    mu = f32(quantization_channels - 1)
    audio = np.asarray(audio, dtype=np.float32)
    magnitude = (np.log(f32(1) + mu * np.abs(audio)).astype(np.float32)
                 / np.log(f32(1.0) + mu).astype(np.float32)).astype(np.float32)
    signal = (np.sign(audio).astype(np.float32) * magnitude).astype(np.float32)
    out = ((signal + f32(1)) / f32(2) * mu + f32(0.5)).astype(np.float32)
    return out.astype(np.int32)


def mu_law_decode(output, quantization_channels):
    """ops.py:76-85.  float32 throughout; ``1/mu`` is a Python float folded to a
    float32 constant (``from __future__ import division``, ops.py:1)."""
    f32 = np.float32
    mu = quantization_channels - 1
    casted = np.asarray(output).astype(np.float32)
    signal = (f32(2) * (casted / f32(mu)) - f32(1)).astype(np.float32)
    magnitude = (f32(1.0 / mu) *
                 (np.power(f32(1 + mu), np.abs(signal)).astype(np.float32) - f32(1))).astype(np.float32)
    return (np.sign(signal).astype(np.float32) * magnitude).astype(np.float32)


# --------------------------------------------------------------------------- #
# time<->batch reshapes and the dilated causal convolution     ops.py:27-62
# --------------------------------------------------------------------------- #


def time_to_batch(value: torch.Tensor, dilation: int) -> torch.Tensor:
    """ops.py:27-34."""
    b, t, c = value.shape
    pad_elements = dilation - 1 - (t + dilation - 1) % dilation
    padded = F.pad(value, (0, 0, 0, pad_elements))
    reshaped = padded.reshape(-1, dilation, c)
    transposed = reshaped.permute(1, 0, 2)
    return transposed.reshape(b * dilation, -1, c)


def batch_to_time(value: torch.Tensor, dilation: int) -> torch.Tensor:
    """ops.py:37-43."""
    s0, _, c = value.shape
    prepared = value.reshape(dilation, -1, c)
    transposed = prepared.permute(1, 0, 2)
    return transposed.reshape(s0 // dilation, -1, c)


def _conv1d_same(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """tf.nn.conv1d(x[B,T,Cin], filt[W,Cin,Cout], stride=1, padding='SAME').

    TF SAME padding: total = W-1, left = total//2, right = total-left, and the op is a
    cross-correlation: out[t] = sum_k in_padded[t+k] . filt[k]."""
    w = filt.shape[0]
    total = w - 1
    left = total // 2
    right = total - left
    xp = F.pad(x, (0, 0, left, right))
    # torch conv1d wants [B,C,T] and weight [Cout,Cin,W]
    out = F.conv1d(xp.permute(0, 2, 1), filt.permute(2, 1, 0))
    return out.permute(0, 2, 1)


def causal_conv(value: torch.Tensor, filter_: torch.Tensor, dilation: int) -> torch.Tensor:
    """ops.py:46-62, op for op (pad -> time_to_batch -> conv1d SAME -> batch_to_time -> slice)."""
    filter_width = filter_.shape[0]
    padded = F.pad(value, (0, 0, (filter_width - 1) * dilation, 0))
    if dilation > 1:
        transformed = time_to_batch(padded, dilation)
        conv = _conv1d_same(transformed, filter_)
        restored = batch_to_time(conv, dilation)
    else:
        restored = _conv1d_same(padded, filter_)
    return restored[:, :value.shape[1], :]


def causal_conv_closed_form(value: torch.Tensor, filter_: torch.Tensor, dilation: int) -> torch.Tensor:
    """Closed form of causal_conv for filter width 2:
    y[t] = x[t-d].w[0] + x[t].w[1], zeros for t<d, per batch element."""
    assert filter_.shape[0] == 2
    past = F.pad(value, (0, 0, dilation, 0))[:, :value.shape[1], :]
    return past @ filter_[0] + value @ filter_[1]


# --------------------------------------------------------------------------- #
# variables                                        wavenet/model.py:7-28,118-225
# --------------------------------------------------------------------------- #


def xavier_uniform(rng: np.random.Generator, shape):
    """tf.contrib.layers.xavier_initializer_conv2d() (uniform): fan_in/out include the
    receptive field (product of leading dims).  model.py:7-12."""
    recept = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in = shape[-2] * recept
    fan_out = shape[-1] * recept
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def variable_specs(dilations, filter_width, residual_channels, dilation_channels, skip_channels,
                   quantization_channels=256, use_biases=False, scalar_input=False,
                   initial_filter_width=32, global_condition_channels=None,
                   global_condition_cardinality=None):
    """Names (checkpoint keys), shapes and init kind in creation order, model.py:118-225.
    Bias variables are emitted under their *intended* names."""
    specs = []
    if global_condition_cardinality is not None:
        kind = 'identity' if global_condition_cardinality == global_condition_channels else 'xavier'
        specs.append(('wavenet/embeddings/gc_embedding',
                      (global_condition_cardinality, global_condition_channels), kind))
    if scalar_input:
        specs.append(('wavenet/causal_layer/filter', (initial_filter_width, 1, residual_channels), 'xavier'))
    else:
        specs.append(('wavenet/causal_layer/filter',
                      (filter_width, quantization_channels, residual_channels), 'xavier'))
    for i, _ in enumerate(dilations):
        p = 'wavenet/dilated_stack/layer{}/'.format(i)
        specs.append((p + 'filter', (filter_width, residual_channels, dilation_channels), 'xavier'))
        specs.append((p + 'gate', (filter_width, residual_channels, dilation_channels), 'xavier'))
        specs.append((p + 'dense', (1, dilation_channels, residual_channels), 'xavier'))
        specs.append((p + 'skip', (1, dilation_channels, skip_channels), 'xavier'))
        if global_condition_channels is not None:
            specs.append((p + 'gc_gate', (1, global_condition_channels, dilation_channels), 'xavier'))
            specs.append((p + 'gc_filter', (1, global_condition_channels, dilation_channels), 'xavier'))
        if use_biases:
            specs.append((p + 'filter_bias', (dilation_channels,), 'zeros'))
            specs.append((p + 'gate_bias', (dilation_channels,), 'zeros'))
            specs.append((p + 'dense_bias', (residual_channels,), 'zeros'))
            specs.append((p + 'slip_bias', (skip_channels,), 'zeros'))  # sic, model.py:203
    p = 'wavenet/postprocessing/'
    specs.append((p + 'postprocess1', (1, skip_channels, skip_channels), 'xavier'))
    specs.append((p + 'postprocess2', (1, skip_channels, quantization_channels), 'xavier'))
    if use_biases:
        specs.append((p + 'postprocess1_bias', (skip_channels,), 'zeros'))
        specs.append((p + 'postprocess2_bias', (quantization_channels,), 'zeros'))
    return specs


def init_variables(specs, seed=0, bias_scale=0.0):
    """Seeded initial values.  ``bias_scale`` > 0 draws non-zero biases so that parity
    tests exercise the bias paths (the reference initialises biases to 0, model.py:24-28)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, kind in specs:
        if kind == 'xavier':
            out[name] = xavier_uniform(rng, shape)
        elif kind == 'identity':
            out[name] = np.identity(shape[0], dtype=np.float32)
        else:
            if bias_scale > 0:
                out[name] = (bias_scale * rng.standard_normal(shape)).astype(np.float32)
            else:
                out[name] = np.zeros(shape, dtype=np.float32)
    return out


# --------------------------------------------------------------------------- #
# TF32 operand-rounding emulation (arithmetic-matched variant of the oracle)
# --------------------------------------------------------------------------- #
# The sm_100a kernels feed the tensor cores TF32 operands (fp32 rounded to a 10-bit mantissa,
# round-to-nearest ties-away = PTX cvt.rna.tf32.f32) and accumulate in fp32.  `emulate='tf32'`
# makes the oracle apply the same rounding at the same operand positions, so that kernel LOGIC
# can be checked tightly; parity against the un-rounded oracle is asserted separately at the
# tolerance BASELINE.json states.


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    x32 = x.detach().to(torch.float32).contiguous()
    i = x32.view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    return r.to(x.dtype)


class _RoundedMatmul(torch.autograd.Function):
    """y = a @ b with TF32-rounded operands.  fwd_round=False keeps the forward product exact
    (the forward block kernel uses a 3-term split that is fp32-grade) while the backward
    products still see rounded operands (single-pass TF32 gradient kernels)."""

    @staticmethod
    def forward(ctx, a, b, fwd_round):
        ctx.save_for_backward(a, b)
        if fwd_round:
            return round_tf32(a) @ round_tf32(b)
        return a @ b

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        gr = round_tf32(g)
        ga = gr @ round_tf32(b).transpose(-1, -2)
        gb = round_tf32(a).reshape(-1, a.shape[-1]).t() @ gr.reshape(-1, g.shape[-1])
        return ga, gb, None


# --------------------------------------------------------------------------- #
# softmax cross entropy with TF-0.10 forward AND backward semantics
# --------------------------------------------------------------------------- #


class _TFSoftmaxXent(torch.autograd.Function):
    """tf.nn.softmax_cross_entropy_with_logits(logits, labels) as used at model.py:663-665.

    forward : loss_i = -sum_j labels_ij * log_softmax(logits)_ij
    backward: TF's kernel emits ``backprop = softmax - labels`` (xent_op: it assumes each
              label row sums to one) and the registered gradient multiplies by the incoming
              grad.  For the all-zero label row at the last time step (model.py:657-659)
              the loss is exactly 0 but the gradient is softmax(logits), not 0."""

    @staticmethod
    def forward(ctx, logits, labels):
        logp = torch.log_softmax(logits, dim=-1)
        ctx.save_for_backward(logp, labels)
        return -(labels * logp).sum(dim=-1)

    @staticmethod
    def backward(ctx, grad_out):
        logp, labels = ctx.saved_tensors
        return grad_out.unsqueeze(-1) * (logp.exp() - labels), None


# --------------------------------------------------------------------------- #
# the model                                                wavenet/model.py
# --------------------------------------------------------------------------- #


class OracleWaveNet(object):
    """CPU restatement of ``WaveNetModel`` (model.py:31-685) with injected weights."""

    def __init__(self, batch_size, dilations, filter_width, residual_channels, dilation_channels,
                 skip_channels, quantization_channels=2 ** 8, use_biases=False, scalar_input=False,
                 initial_filter_width=32, histograms=False, global_condition_channels=None,
                 global_condition_cardinality=None, residual_postproc=False,
                 dtype=torch.float32, seed=0, bias_scale=0.0, faithful=True, emulate=None):
        self.batch_size = batch_size
        self.dilations = list(dilations)
        self.filter_width = filter_width
        self.residual_channels = residual_channels
        self.dilation_channels = dilation_channels
        self.skip_channels = skip_channels
        self.quantization_channels = quantization_channels
        self.use_biases = use_biases
        self.scalar_input = scalar_input
        self.initial_filter_width = initial_filter_width
        self.global_condition_channels = global_condition_channels
        self.global_condition_cardinality = global_condition_cardinality
        self.residual_postproc = residual_postproc
        self.dtype = dtype
        self.faithful = faithful  # True: reshape-chain causal_conv; False: closed form
        self.emulate = emulate    # None | 'tf32' (closed-form path only): see round_tf32 above
        assert emulate in (None, 'tf32') and not (emulate and faithful)
        self.specs = variable_specs(dilations, filter_width, residual_channels, dilation_channels,
                                    skip_channels, quantization_channels, use_biases, scalar_input,
                                    initial_filter_width, global_condition_channels,
                                    global_condition_cardinality)
        self.load_state_dict(init_variables(self.specs, seed=seed, bias_scale=bias_scale))
        self._gen_state = None

    # -- state ---------------------------------------------------------------
    def load_state_dict(self, sd):
        self.vars = {}
        for name, shape, _ in self.specs:
            a = np.asarray(sd[name], dtype=np.float32)
            assert tuple(a.shape) == tuple(shape), (name, a.shape, shape)
            self.vars[name] = torch.tensor(a, dtype=self.dtype, requires_grad=True)

    def state_dict(self):
        return {k: v.detach().to(torch.float32).numpy().copy() for k, v in self.vars.items()}

    def _lv(self, i, key):
        return self.vars['wavenet/dilated_stack/layer{}/{}'.format(i, key)]

    def _cc(self, x, w, d):
        if self.faithful or w.shape[0] != 2:
            return causal_conv(x, w, d)
        if self.emulate:
            past = F.pad(x, (0, 0, d, 0))[:, :x.shape[1], :]
            return _RoundedMatmul.apply(past, w[0], False) + _RoundedMatmul.apply(x, w[1], False)
        return causal_conv_closed_form(x, w, d)

    def _conv1x1(self, x, w, kind):
        """tf.nn.conv1d with a width-1 filter; kind: 'block' (dense) or 'gemm' (skip/post)."""
        if self.emulate:
            return _RoundedMatmul.apply(x, w[0], kind == 'gemm')
        return _conv1d_same(x, w)

    # -- pieces ----------------------------------------------------------------
    def _one_hot(self, ids):
        """model.py:518-531.  Out-of-range ids give an all-zero row (tf.one_hot)."""
        ids = torch.as_tensor(np.asarray(ids), dtype=torch.int64)
        q = self.quantization_channels
        valid = (ids >= 0) & (ids < q)
        oh = F.one_hot(ids.clamp(0, q - 1), q).to(self.dtype) * valid.unsqueeze(-1).to(self.dtype)
        return oh.reshape(self.batch_size, -1, q)

    def _embed_gc(self, global_condition):
        """model.py:533-562 (integer-category branch; the dense-vector branch is passed through)."""
        if global_condition is None:
            return None
        g = self.global_condition_channels
        if self.global_condition_cardinality is not None:
            ids = torch.as_tensor(np.asarray(global_condition), dtype=torch.int64).reshape(-1)
            emb = self.vars['wavenet/embeddings/gc_embedding'][ids]
        else:
            emb = torch.as_tensor(np.asarray(global_condition), dtype=self.dtype)
        return emb.reshape(self.batch_size, 1, g)

    def _dilation_layer(self, x, i, d, gc, is_last):
        """model.py:236-330."""
        conv_filter = self._cc(x, self._lv(i, 'filter'), d)
        conv_gate = self._cc(x, self._lv(i, 'gate'), d)
        if gc is not None:
            conv_filter = conv_filter + _conv1d_same(gc, self._lv(i, 'gc_filter'))
            conv_gate = conv_gate + _conv1d_same(gc, self._lv(i, 'gc_gate'))
        if self.use_biases:
            conv_filter = conv_filter + self._lv(i, 'filter_bias')
            conv_gate = conv_gate + self._lv(i, 'gate_bias')
        out = torch.tanh(conv_filter) * torch.sigmoid(conv_gate)
        if self.emulate:
            # the kernel stores z tf32-rounded (straight-through for the gradient)
            out = out + (round_tf32(out) - out).detach()
        transformed = None
        if not is_last:
            transformed = self._conv1x1(out, self._lv(i, 'dense'), 'block')
        skip = self._conv1x1(out, self._lv(i, 'skip'), 'gemm')
        if self.use_biases:
            if not is_last:
                transformed = transformed + self._lv(i, 'dense_bias')
            skip = skip + self._lv(i, 'slip_bias')
        if is_last:
            return skip, None
        return skip, x + transformed

    def network(self, input_batch, gc, return_intermediates=False):
        """model.py:389-442."""
        cur = self._cc(input_batch, self.vars['wavenet/causal_layer/filter'], 1)
        outputs = []
        xs = [cur]
        n = len(self.dilations)
        for i, d in enumerate(self.dilations):
            skip, cur = self._dilation_layer(cur, i, d, gc, i == n - 1)
            outputs.append(skip)
            if cur is not None:
                xs.append(cur)
        p = 'wavenet/postprocessing/'
        total = sum(outputs)
        t1 = torch.relu(total)
        conv1 = self._conv1x1(t1, self.vars[p + 'postprocess1'], 'gemm')
        if self.use_biases:
            conv1 = conv1 + self.vars[p + 'postprocess1_bias']
        t2 = torch.relu(conv1)
        if self.residual_postproc:
            t2 = t2 + total
        conv2 = self._conv1x1(t2, self.vars[p + 'postprocess2'], 'gemm')
        if self.use_biases:
            conv2 = conv2 + self.vars[p + 'postprocess2_bias']
        if return_intermediates:
            return conv2, dict(xs=xs, total=total, conv1=conv1)
        return conv2

    # -- public API ---------------------------------------------------------
    def loss(self, input_batch, global_condition_batch=None, l2_regularization_strength=None,
             return_logits=False):
        """model.py:628-685."""
        q = self.quantization_channels
        audio = np.asarray(input_batch, dtype=np.float32)
        encoded_input = mu_law_encode(audio, q)
        gc = self._embed_gc(global_condition_batch)
        encoded = self._one_hot(encoded_input)
        if self.scalar_input:
            net_in = torch.as_tensor(audio, dtype=self.dtype).reshape(self.batch_size, -1, 1)
        else:
            net_in = encoded
        raw = self.network(net_in, gc)
        shifted = F.pad(encoded[:, 1:, :], (0, 0, 0, 1))
        per = _TFSoftmaxXent.apply(raw.reshape(-1, q), shifted.reshape(-1, q))
        reduced = per.mean()
        if l2_regularization_strength is not None:
            # model.py:674-676: the name filter never matches (App. B naming bug) so biases
            # are included in this snapshot.
            l2 = sum((v * v).sum() / 2 for v in self.vars.values())
            reduced = reduced + l2_regularization_strength * l2
        if return_logits:
            return reduced, raw
        return reduced

    def predict_proba(self, waveform, global_condition=None):
        """model.py:564-590: full network on the window, float64 softmax, last row."""
        q = self.quantization_channels
        with torch.no_grad():
            if self.scalar_input:
                enc = torch.as_tensor(mu_law_decode(np.asarray(waveform), q), dtype=self.dtype)
                enc = enc.reshape(self.batch_size, -1, 1)
            else:
                enc = self._one_hot(waveform)
            gc = self._embed_gc(global_condition)
            raw = self.network(enc, gc).reshape(-1, q)
            proba = torch.softmax(raw.to(torch.float64), dim=-1).to(torch.float32)
            return proba[-1].numpy()

    # -- incremental generator                                model.py:332-387,444-516,592-626
    def init_ops(self):
        """model.py:457-463,477-484: queues pre-filled with zeros."""
        b = self.batch_size
        self._gen_state = dict(
            causal=torch.zeros(1, b, self.quantization_channels, dtype=self.dtype),
            layers=[torch.zeros(d, b, self.residual_channels, dtype=self.dtype) for d in self.dilations],
            pos=[0 for _ in self.dilations])

    def predict_proba_incremental(self, waveform, global_condition=None, push=True, return_logits=False):
        """One step of model.py:592-626 with ``push`` standing for fetching ``push_ops``.

        Returns the last batch row like the reference (model.py:622-626) unless
        ``return_logits`` (then [B,Q] raw logits)."""
        if self.filter_width > 2:
            raise NotImplementedError("Incremental generation does not support filter_width > 2.")
        if self.scalar_input:
            raise NotImplementedError("Scalar input is not supported by fast generation.")
        if self._gen_state is None:
            self.init_ops()
        st = self._gen_state
        q = self.quantization_channels
        with torch.no_grad():
            gc = self._embed_gc(global_condition)
            cur = self._one_hot(waveform).reshape(-1, q)
            # causal layer: FIFO of capacity 1
            state = st['causal'][0]
            w = self.vars['wavenet/causal_layer/filter']
            new_causal = cur
            cur = state @ w[0] + cur @ w[1]
            outputs = []
            pushes = []
            for i, d in enumerate(self.dilations):
                pos = st['pos'][i]
                state = st['layers'][i][pos]          # dequeue: element pushed d steps ago
                pushes.append(cur)
                wf, wg = self._lv(i, 'filter'), self._lv(i, 'gate')
                of = state @ wf[0] + cur @ wf[1]
                og = state @ wg[0] + cur @ wg[1]
                if gc is not None:
                    g = gc.reshape(1, -1)             # model.py:360-361 (B=1 only)
                    of = of + g @ self._lv(i, 'gc_filter')[0]
                    og = og + g @ self._lv(i, 'gc_gate')[0]
                if self.use_biases:
                    of = of + self._lv(i, 'filter_bias')
                    og = og + self._lv(i, 'gate_bias')
                out = torch.tanh(of) * torch.sigmoid(og)
                transformed = out @ self._lv(i, 'dense')[0]
                if self.use_biases:
                    transformed = transformed + self._lv(i, 'dense_bias')
                skip = out @ self._lv(i, 'skip')[0]
                if self.use_biases:
                    skip = skip + self._lv(i, 'slip_bias')
                outputs.append(skip)
                cur = cur + transformed
            p = 'wavenet/postprocessing/'
            total = sum(outputs)
            t1 = torch.relu(total)
            conv1 = t1 @ self.vars[p + 'postprocess1'][0]
            if self.use_biases:
                conv1 = conv1 + self.vars[p + 'postprocess1_bias']
            t2 = torch.relu(conv1)                      # no residual_postproc here (model.py:505-514)
            conv2 = t2 @ self.vars[p + 'postprocess2'][0]
            if self.use_biases:
                conv2 = conv2 + self.vars[p + 'postprocess2_bias']
            if push:
                st['causal'][0] = new_causal
                for i, d in enumerate(self.dilations):
                    st['layers'][i][st['pos'][i]] = pushes[i]
                    st['pos'][i] = (st['pos'][i] + 1) % d
            if return_logits:
                return conv2.to(torch.float32).numpy()
            out = conv2.reshape(-1, q)
            proba = torch.softmax(out.to(torch.float64), dim=-1).to(torch.float32)
            return proba[-1].numpy()

    # -- gradient helper -------------------------------------------------------
    def loss_and_grads(self, input_batch, global_condition_batch=None, l2=None):
        for v in self.vars.values():
            v.grad = None
        loss, logits = self.loss(input_batch, global_condition_batch, l2, return_logits=True)
        loss.backward()
        grads = {}
        for k, v in self.vars.items():
            grads[k] = (v.grad.detach().to(torch.float32).numpy().copy() if v.grad is not None
                        else np.zeros(tuple(v.shape), np.float32))
        return float(loss.detach()), logits.detach().to(torch.float32).numpy(), grads


# --------------------------------------------------------------------------- #
# optimizers with TF-0.10 update rules                wavenet/ops.py:6-24
# --------------------------------------------------------------------------- #


class TFOptimizer(object):
    def __init__(self, kind, learning_rate, momentum):
        self.kind, self.lr, self.mu = kind, float(learning_rate), float(momentum)
        self.t = 0
        self.slots = {}

    def apply(self, params: dict, grads: dict):
        """In-place update of float32 numpy ``params`` with ``grads`` (None grad = skipped var)."""
        self.t += 1
        f32 = np.float32
        for k, w in params.items():
            g = grads.get(k)
            if g is None:
                continue
            g = g.astype(np.float32)
            if self.kind == 'sgd':       # MomentumOptimizer: a = mu*a + g ; w -= lr*a
                a = self.slots.setdefault(k, np.zeros_like(w))
                a[...] = f32(self.mu) * a + g
                w -= f32(self.lr) * a
            elif self.kind == 'adam':    # AdamOptimizer(lr, eps=1e-4), beta1=.9, beta2=.999
                m, v = self.slots.setdefault(k, (np.zeros_like(w), np.zeros_like(w)))
                b1, b2, eps = 0.9, 0.999, 1e-4
                lr_t = self.lr * math.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
                m[...] = f32(b1) * m + f32(1 - b1) * g
                v[...] = f32(b2) * v + f32(1 - b2) * g * g
                w -= f32(lr_t) * m / (np.sqrt(v) + f32(eps))
            elif self.kind == 'rmsprop':  # RMSPropOptimizer(lr, decay=.9, momentum, eps=1e-5); ms init ones
                ms, mom = self.slots.setdefault(k, (np.ones_like(w), np.zeros_like(w)))
                ms[...] = f32(0.9) * ms + f32(0.1) * g * g
                mom[...] = f32(self.mu) * mom + f32(self.lr) * g / np.sqrt(ms + f32(1e-5))
                w -= mom
            else:
                raise KeyError(self.kind)


# --------------------------------------------------------------------------- #
# sampling arithmetic of the generation loop               generate.py:228-241
# --------------------------------------------------------------------------- #


def scale_prediction(prediction, temperature):
    """generate.py:229-233 (float32 numpy)."""
    prediction = np.asarray(prediction, dtype=np.float32)
    with np.errstate(divide='ignore'):
        scaled = (np.log(prediction) / np.float32(temperature)).astype(np.float32)
        scaled = scaled - np.logaddexp.reduce(scaled)
        return np.exp(scaled).astype(np.float32)


def choice_from_uniform(p, u):
    """``np.random.choice(np.arange(Q), p=p)`` (generate.py:239-240) given the uniform
    double ``u`` that RandomState.random_sample() would have produced: the legacy
    algorithm is cdf = cumsum(float64(p)); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side='right'))
