/* wavenet_b200.h -- C ABI of libwavenet_b200.so (sm_100a).
 *
 * Drop-in boundary for the dilated-causal-convolution hot path of
 * jyegerlehner/tensorflow-wavenet.  The reference has no FFI of its own: its boundary is the
 * Python API of wavenet/model.py + wavenet/ops.py, whose arithmetic is executed by stock
 * TensorFlow ops.  Every entry point below replaces the TF library-op call sites of one
 * reference function (cited as file:line relative to the reference root); the Python mirror
 * of the reference API (tensorflow-wavenet_b200/wavenet/) binds them with ctypes -- see
 * INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a CALLER-OWNED DEVICE pointer unless it says "host"; activations are
 *     row-major [B*T, C] (time-major within a batch element, channels contiguous), float32;
 *   - every call is asynchronous on `stream` (a cudaStream_t) and allocates nothing;
 *   - return value: 0 ok, < 0 bad argument / unsupported shape, > 0 a cudaError_t;
 *   - ONE DEVICE PER PROCESS (the deployment model: one process per GPU, torch.distributed / NCCL between them): the
 *     library's side streams, fork / join events, cached function attributes and SM count are process-global and
 *     belong to the device that was current at the first call; calls are not thread-safe.
 */
#ifndef WAVENET_B200_H_
#define WAVENET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WN_MAX_LAYERS 128
#define WN_ABI_VERSION 2

typedef void* wn_stream_t; /* cudaStream_t */

/* Hyper-parameters of WaveNetModel.__init__ (wavenet/model.py:46-116) that shape the hot path. */
typedef struct wn_config {
  int32_t n_layers;              /* len(dilations)                         */
  int32_t residual_channels;     /* R                                      */
  int32_t dilation_channels;     /* D  (fast generation needs D == R in {16, 32}; training takes any widths) */
  int32_t skip_channels;         /* S                                      */
  int32_t quantization_channels; /* Q                                      */
  int32_t gc_channels;           /* G, 0 = no global conditioning          */
  int32_t gc_cardinality;        /* rows of the embedding table, 0 = none  */
  int32_t use_biases;
  int32_t residual_postproc;
  int32_t dilations[WN_MAX_LAYERS];
  /* ABI 2: scalar_input front end (model.py:143-153): the causal layer is a width-initial_filter_width convolution of the
   * raw float waveform ([IFW, 1, R] filter) instead of a width-2 convolution of the one-hot encoding ([2, Q, R]) */
  int32_t scalar_input;
  int32_t initial_filter_width;
} wn_config;

/* Offsets (in floats) of each variable group inside the flat parameter / gradient buffer.
 * Groups are stored layer-major so that e.g. all skip weights form the [L*D, S] operand of
 * the skip-sum GEMM.  Shapes are the reference's (model.py:118-225). */
typedef struct wn_layout {
  int64_t causal;        /* [2, Q, R] ([IFW, 1, R] with scalar_input)   wavenet/causal_layer/filter */
  int64_t filter;        /* [L][2, R, D]         .../layer{i}/filter                     */
  int64_t gate;          /* [L][2, R, D]         .../layer{i}/gate                       */
  int64_t dense;         /* [L][D, R]            .../layer{i}/dense                      */
  int64_t skip;          /* [L][D, S]            .../layer{i}/skip                       */
  int64_t gc_filter;     /* [L][G, D]            .../layer{i}/gc_filter        (-1: none) */
  int64_t gc_gate;       /* [L][G, D]            .../layer{i}/gc_gate          (-1: none) */
  int64_t filter_bias;   /* [L][D]                                             (-1: none) */
  int64_t gate_bias;     /* [L][D]                                             (-1: none) */
  int64_t dense_bias;    /* [L][R]                                             (-1: none) */
  int64_t skip_bias;     /* [L][S]               ("slip_bias", model.py:203)   (-1: none) */
  int64_t post1;         /* [S, S]               postprocessing/postprocess1             */
  int64_t post2;         /* [S, Q]               postprocessing/postprocess2             */
  int64_t post1_bias;    /* [S]                                                (-1: none) */
  int64_t post2_bias;    /* [Q]                                                (-1: none) */
  int64_t gc_embedding;  /* [card, G]            embeddings/gc_embedding       (-1: none) */
  int64_t total;         /* number of floats in the flat buffer                          */
} wn_layout;

int wn_abi_version(void);
/* Validates cfg (returns <0 when this build has no kernel for it) and fills `out`. */
int wn_param_layout(const wn_config* cfg, wn_layout* out);

/* ---- mu-law companding: wavenet/ops.py:65-73 (encode), :76-85 (decode) ------------------
 * thresholds: Q-1 float32 decision levels, lut: Q float32 levels, both built by the host
 * mirror from the float32 formulas (bit-exact integer result by construction). */
int wn_mulaw_encode(const float* audio, int64_t n, const float* thresholds, int32_t q, int32_t* ids,
                    wn_stream_t stream);
int wn_mulaw_decode(const int32_t* ids, int64_t n, const float* lut, int32_t q, float* out,
                    wn_stream_t stream);

/* ---- one-hot + causal layer: model.py:518-531 + :227-234 (ops.py:46-62 with dilation 1) --- */
int wn_frontend_fwd(const int32_t* ids, const float* causal_filter, float* x0, int32_t batch, int32_t time,
                    int32_t q, int32_t r, wn_stream_t stream);
int wn_frontend_bwd(const int32_t* ids, const float* dx0, float* grad_causal_filter, int32_t batch,
                    int32_t time, int32_t q, int32_t r, wn_stream_t stream);

/* ---- generic width-2 dilated causal convolution: ops.py:46-62 (causal_conv) ----------------
 * y[b,t,:] = x[b,t-d,:] . w[0] + x[b,t,:] . w[1]   (fp32 FMA path; the public ops.causal_conv) */
int wn_causal_conv(const float* x, const float* w, float* y, int32_t batch, int32_t time, int32_t cin,
                   int32_t cout, int32_t width, int32_t dilation, wn_stream_t stream);

/* ---- one gated residual block: model.py:236-330 (_create_dilation_layer) -------------------
 * prebias: [B, 2D] = [filter_bias | gate_bias] + gc projection (wn_cond_bias_fwd), never NULL.
 * z is written (tf32-rounded) to zcat + row*ldz; x_out is not written when is_last.          */
int wn_block_fwd(const float* x, float* x_out, float* zcat, int32_t ldz, const float* filter,
                 const float* gate, const float* dense, const float* prebias, const float* dense_bias,
                 int32_t batch, int32_t time, int32_t dilation, int32_t channels, int32_t is_last,
                 wn_stream_t stream);
/* Backward of the block (TF autodiff of the same lines).  dz_skip is the gradient wrt z coming
 * from the skip GEMM (row stride ldz); dpre_scratch is [B*T, 2D]; gradient outputs accumulate. */
int wn_block_bwd(const float* x, const float* dx_out, const float* dz_skip, int32_t ldz, float* dx,
                 float* dpre_scratch, const float* zcat, const float* filter, const float* gate,
                 const float* dense, const float* prebias, float* grad_filter, float* grad_gate,
                 float* grad_dense, float* grad_prebias, float* grad_dense_bias, int32_t batch,
                 int32_t time, int32_t dilation, int32_t channels, int32_t is_last, wn_stream_t stream);

/* ---- 1x1 convolutions as GEMMs (TF32 tensor cores, fp32 accumulate): model.py:304-305,432,438
 * mode 0: C[M,N] = A[M,K].B[K,N]; 1: C = A[M,K].B[N,K]^T; 2: C += A[K,M]^T.B[K,N] (atomic).
 * flags: 1 relu, 2 round output to tf32, 4 atomic accumulate. */
int wn_gemm_tf32(int32_t mode, const float* a, int32_t lda, const float* b, int32_t ldb, float* c,
                 int32_t ldc, int32_t m, int32_t n, int32_t k, const float* bias, const float* relu_mask,
                 int32_t ldmask, int32_t flags, int32_t split_k, wn_stream_t stream);

/* The production GEMM: C[M,N] (+)= A[M,K] . B[N,K]^T, both operands K-major, on tcgen05 (TMA +
 * tcgen05.mma.kind::tf32 + TMEM).  ct (nullable) receives a transposed copy [N][M] (row pitch ldct).
 * flags as above; flag 4 (atomic) enables split-K over `split_k` CTAs along K. */
int wn_gemm_nt_umma(const float* a, int32_t lda, const float* b, int32_t ldb, float* c, int32_t ldc, float* ct,
                    int32_t ldct, int32_t m, int32_t n, int32_t k, const float* bias, const float* relu_mask,
                    int32_t ldmask, int32_t flags, int32_t split_k, wn_stream_t stream);

/* The same tcgen05 kernel in all three operand forms of wn_gemm_tf32 (mode 0 NN / 1 NT / 2 TN), every
 * operand read as it lies in HBM (MN-major operands through the 32-byte-atom 128B swizzle).  Returns -3
 * for shapes it does not take (N % 32 for modes 0/2, M % 32 for mode 2, leading dimensions % 4). */
int wn_gemm_umma(int32_t mode, const float* a, int32_t lda, const float* b, int32_t ldb, float* c, int32_t ldc,
                 int32_t m, int32_t n, int32_t k, const float* bias, const float* relu_mask, int32_t ldmask,
                 int32_t flags, int32_t split_k, wn_stream_t stream);

/* The fp16 form of the post-processing products (model.py:430-440 forward, and their input gradients):
 * v[m,n] = mask(relu(a16[m,k] . b16[n,k]^T + bias)), both operands IEEE fp16 with k contiguous, fp32
 * accumulation; flags: 1 relu, 2 tf32 rounding of c; relu_mask (optional, fp32 [m,n]) zeroes v where mask <= 0.
 * c = c_scale * v as fp32; c16 (optional) = v rounded to fp16 (NOT scaled) -- the next product's operand.
 * The gradient chain runs in a domain scaled by a power of two: c16 stays in it, c_scale takes c out of it.
 * lda/ldb/ldc16 multiples of 8, ldc/ldmask multiples of 4; returns -3 otherwise. */
int wn_gemm_f16_nt(const void* a16, int32_t lda, const void* b16, int32_t ldb, float* c, int32_t ldc, void* c16,
                   int32_t ldc16, int32_t m, int32_t n, int32_t k, const float* bias, const float* relu_mask,
                   int32_t ldmask, float c_scale, int32_t flags, wn_stream_t stream);

/* The same product where only the fp16 copy is wanted (c may be null) and the epilogue also ADDS colsum_scale * (column
 * sums of c16 over all m rows) to colsum[n]: how the training step gets the bias gradient of the layer below out of the
 * input-gradient GEMM that produces the matrix (autodiff of model.py:432-440).  No bias argument in this form. */
int wn_gemm_f16_nt_colsum(const void* a16, int32_t lda, const void* b16, int32_t ldb, float* c, int32_t ldc, void* c16,
                          int32_t ldc16, int32_t m, int32_t n, int32_t k, float c_scale, float* colsum, float colsum_scale,
                          wn_stream_t stream);

/* Weight-gradient form on fp16 operands: c[m,n] += c_scale * sum_k a16[k,m] * b16[k,n]  (a16 [k][lda], b16 [k][ldb] as
 * they lie in memory: time is the row index; fp32 accumulation, split over k, red.global.add into c).
 * m, n multiples of 64, lda/ldb multiples of 8; returns -3 otherwise. */
int wn_gemm_f16_tn(const void* a16, int32_t lda, const void* b16, int32_t ldb, float* c, int32_t ldc, int32_t m, int32_t n,
                   int32_t k, float c_scale, int32_t split_k, wn_stream_t stream);

/* ---- softmax cross entropy vs the next sample: model.py:654-666 ---------------------------
 * logits [B*T, Q] are overwritten by d loss / d logits (TF backprop semantics) when write_grad. */
int wn_softmax_xent(float* logits, const int32_t* ids, int32_t batch, int32_t time, int32_t q,
                    float* partials, int32_t n_partials, float* loss_out, int32_t write_grad,
                    wn_stream_t stream);

/* postprocess2 + softmax cross entropy in one kernel (model.py:438-440 + 654-666), the form the training step runs when
 * quantization_channels == 256: a16 [B*T, k] fp16 (relu(conv1) rows), w16 [256, k] fp16 (postprocess2 transposed),
 * bias [256] fp32 or null.  Writes loss (mean over B*T rows), g16 [B*T, 256] fp16 = (softmax - onehot) * grad_scale and,
 * when bias_grad is given, ADDS colsum_scale * column sums of g16 to bias_grad[256].  partials: >= 148 floats.
 * The logits themselves are never stored.  Returns -2 unless q == 256 and k is a multiple of 64. */
int wn_post2_xent(const void* a16, int32_t lda, const void* w16, int32_t ldw, const float* bias, const int32_t* ids,
                  int32_t batch, int32_t time, int32_t k, int32_t q, float* partials, float* loss_out, void* g16,
                  float grad_scale, float* bias_grad, float colsum_scale, wn_stream_t stream);

/* WaveNetModel.predict_proba (model.py:564-590): the network over the whole window, post-processing on its LAST row only,
 * float64 softmax cast to float32.  proba: [Q].  ids as in wn_forward_logits; workspace: wn_forward_workspace_bytes. */
int wn_predict_last(const wn_config* cfg, const float* params, void* workspace, int64_t workspace_bytes, const int32_t* ids,
                    const int32_t* gc_ids, int32_t batch, int32_t time, float* proba, wn_stream_t stream);

/* model.py:670-680: *loss += coef * sum_v tf.nn.l2_loss(v) = coef * sum(params^2) / 2 over the flat parameter buffer (its
 * alignment gaps are zero; in this snapshot of the reference the biases are NOT excluded). */
int wn_add_l2(float* loss, const float* params, int64_t n, float coef, wn_stream_t stream);

/* Data-parallel training (the reference is single-session, train.py:261): `cuda_event` (a cudaEvent_t,
 * NULL = off) is recorded inside every following wn_loss_grad call at the point where the gradients of the TAIL of the flat
 * buffer -- [layout.skip, layout.total): skip, skip_bias, postprocess1/2 and their biases, 80 % of the bytes -- are final,
 * more than a millisecond before the call's last kernel.  A caller that all-reduces that range on another stream behind this
 * event overlaps it with the residual-block backward (wavenet/train_step.py). */
int wn_set_grad_ready_event(void* cuda_event);

/* ---- whole training graph of WaveNetModel.loss: model.py:628-685 + train.py:252 gradients ----
 * audio [B,T] float32 -> loss (device scalar) and the flat gradient buffer (overwritten).
 * workspace: wn_train_workspace_bytes(cfg,batch,time) bytes, 256-byte aligned.               */
int64_t wn_train_workspace_bytes(const wn_config* cfg, int32_t batch, int32_t time);
int wn_loss_grad(const wn_config* cfg, const float* params, float* grads, void* workspace,
                 int64_t workspace_bytes, const float* audio, const int32_t* gc_ids,
                 const float* mulaw_thresholds, int32_t batch, int32_t time, float* loss_out,
                 wn_stream_t stream);
/* Forward only (model.py:389-442 _create_network on encoded ids): logits [B*T, Q] into `logits`. */
int64_t wn_forward_workspace_bytes(const wn_config* cfg, int32_t batch, int32_t time);
int wn_forward_logits(const wn_config* cfg, const float* params, void* workspace, int64_t workspace_bytes,
                      const int32_t* ids, const int32_t* gc_ids, int32_t batch, int32_t time, float* logits,
                      wn_stream_t stream);

/* ---- optimizers, TF-0.10 update rules: wavenet/ops.py:6-24 ---------------------------------
 * grad_scale multiplies the gradient first (1/world_size after a sum all-reduce); l2 adds
 * l2*w to it (model.py:670-680).  step is the 1-based Adam time step. */
int wn_optim_adam(float* w, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, float l2, float grad_scale, wn_stream_t stream);
int wn_optim_momentum(float* w, const float* g, float* accum, int64_t n, double lr, double momentum, float l2,
                      float grad_scale, wn_stream_t stream);
int wn_optim_rmsprop(float* w, const float* g, float* ms, float* mom, int64_t n, double lr, double decay,
                     double momentum, double eps, float l2, float grad_scale, wn_stream_t stream);

/* ---- fast generation: model.py:332-387,444-516,592-626 + generate.py:213-241 ----------------
 * A persistent kernel runs `n_steps` sample steps for `streams` independent streams: per-layer
 * delay lines replace the tf.FIFOQueues, the float64 softmax, temperature scaling and the
 * np.random.choice inverse-cdf draw run in the same kernel, the drawn sample is fed back.
 * state: wn_gen_state_bytes() bytes of device memory, zeroed by wn_gen_reset (init_ops).
 * inputs[streams] / samples_out[streams, n_steps] int32; uniforms[streams, n_steps] float64
 * (host RNG draws; NULL = teacher forcing: step s consumes forced[streams, n_steps] and no
 * sampling happens).  proba_out (nullable) receives the float32 distribution of the LAST step
 * [streams, Q]; commit=0 leaves the delay lines untouched (forward without push_ops). */
int64_t wn_gen_state_bytes(const wn_config* cfg, int32_t streams);
int wn_gen_reset(const wn_config* cfg, void* state, int32_t streams, wn_stream_t stream);
int wn_gen_run(const wn_config* cfg, const float* params, void* state, int32_t streams,
               const int32_t* inputs, const int32_t* forced, const int32_t* gc_ids, const double* uniforms,
               int32_t n_steps, float temperature, int32_t commit, int32_t* samples_out, float* proba_out,
               wn_stream_t stream);
/* Enqueue what the last commit=0, n_steps=1 call computed (push_ops fetched separately). */
int wn_gen_commit(const wn_config* cfg, void* state, int32_t streams, wn_stream_t stream);
/* np.random.choice(arange(Q), p=p) given the uniform double it would draw (generate.py:239-240). */
int wn_sample(const float* proba, const double* uniforms, int32_t rows, int32_t q, int32_t* out,
              wn_stream_t stream);

/* ---- measurement aid (bench.py roofline): CUDA events on the launching stream after every kernel
 * of wn_loss_grad between begin/end; results are summed per kernel kind (wn_profile_tag_name). */
/* Validation switch: run the GEMMs / residual blocks on the legacy mma.sync kernels instead of tcgen05
 * (same arithmetic contract) so that one implementation can be checked against the other.  Workspace
 * sizes depend on the choice: query them again after switching. */
int wn_debug_set_impl(int32_t gemm_mma, int32_t block_mma);
/* fast generation: 1 = allow the latency-mode kernel (one stream over the whole GPU, default), 0 = always the
 * throughput-mode kernel (validation of one against the other). */
int wn_debug_set_gen_impl(int32_t latency_kernel);
/* debug: device buffer of 48 int64 receiving clock64() stamps (4 tiles x 8 phases) of CTA 0, then
 * %globaltimer entry / first-tile / exit stamps of the first, middle and last CTA, of the next
 * wn_block_fwd launches on the tcgen05 path; null disables. */
int wn_debug_timeline(long long* stamps);
/* debug: 8 host-mapped (cudaHostAlloc mapped / pinned) words that receive the identity of the first bounded
 * mbarrier wait that times out before the kernel traps: {barrier smem address, parity, blockDim.x, gridDim.x,
 * blockIdx.x, threadIdx.x, gridDim.y, claimed}. */
int wn_debug_trap_info(unsigned int* host_mapped_words);
#define WN_PROFILE_TAGS 23
int wn_profile_begin(void);
/* records one more profiling event on `stream` (calibration of the per-event overhead: back-to-back marks) */
int wn_profile_mark(int32_t tag, wn_stream_t stream);
int wn_profile_end(float* ms_per_tag /*host*/, int32_t* launches_per_tag /*host*/, int32_t n_tags);
int wn_profile_tag_name(int32_t tag, char* out /*host*/, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* WAVENET_B200_H_ */
