// Internal launcher declarations shared by the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wn {

// ---- optional per-kernel timing (bench.py roofline): CUDA events on the launching stream ----
enum ProfTag {
  PT_MISC = 0, PT_MULAW, PT_COND_BIAS, PT_FRONTEND_FWD, PT_BLOCK_FWD, PT_SKIP_BIAS_SUM, PT_GEMM_SKIP_FWD,
  PT_GEMM_POST1_FWD, PT_GEMM_POST2_FWD, PT_XENT, PT_GEMM_POST2_WGRAD, PT_COLSUM, PT_GEMM_POST2_DGRAD,
  PT_GEMM_POST1_WGRAD, PT_GEMM_POST1_DGRAD, PT_GEMM_SKIP_WGRAD, PT_GEMM_SKIP_DGRAD, PT_BLOCK_BWD_DX,
  PT_BLOCK_WGRAD, PT_FRONTEND_BWD, PT_COND_BIAS_BWD, PT_TRANSPOSE, PT_BLOCK_BWD_PRE, PT_COUNT
};
void prof_mark(cudaStream_t st, int tag);   // no-op unless wn_profile_begin() was called

enum { GEMM_RELU = 1, GEMM_ROUND = 2, GEMM_ATOMIC = 4 };

struct GemmParams {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int M, N, K;
  const float* bias;            // per output column, nullable
  const float* aux; int ldaux;  // relu-gradient mask source (keep where aux > 0), nullable
  float* C2; int ldc2;          // optional second output: acc + bias before relu/mask, nullable
  int flags;
};

int gemm_tf32(int mode, const GemmParams& p, int split_k, cudaStream_t st);
int colsum(const float* A, int lda, int M, int N, float* out, cudaStream_t st);
// scratch (optional, colsum16_scratch_floats(N) floats): per-chunk partial sums + a finishing kernel instead of atomics
int64_t colsum16_scratch_floats(int N);
int colsum16(const void* A16, int lda, int M, int N, float scale, float* out, float* scratch, cudaStream_t st);
// tcgen05 path: C[M,N] (+)= A[M,K] . B[N,K]^T (p.B is the [N,K] operand), optional transposed copy CT[N][M]
int gemm_nt_umma(const GemmParams& p, float* CT, int ldct, int split_k, cudaStream_t st);
// tcgen05 path, mode 0 NN (p.B = [K,N]) / 1 NT (p.B = [N,K]) / 2 TN (p.A = [K,M], p.B = [K,N]; GEMM_ATOMIC)
int gemm_umma(int mode, const GemmParams& p, float* CT, int ldct, int split_k, cudaStream_t st);
bool gemm_umma_supported(int mode, const GemmParams& p);
// extras of gemm_f16_nt for the wide residual blocks (block_wide16.cu)
struct F16Extra {
  int a_split = 0, a_shift = 0;      // two-tap A: k < a_split reads rows m + a_shift, k >= a_split rows m of A16[M][a_split] (K = 2 a_split)
  const void* aux16 = nullptr;       // fp16 [M][N] matrix added (times aux_scale) to the product
  int ldaux16 = 0;
  float aux_scale = 1.f;
  void* gate_z16 = nullptr;          // N = 256 = [f 128 | g 128]: also store z = tanh(f) sigmoid(g) as fp16 [M][128] (row pitch ldz); C16 keeps f | g
  int ldz = 0;
  const void* dpre_P16 = nullptr;    // N = 128: product + aux16 is dz; with the saved [f 128 | g 128] (fp16, row pitch ldp) the epilogue stores
  int ldp = 0;                       // dpre = [df | dg] into C16 ([M][256]); colsum then has 256 entries
};
// forward chain in fp16 (operands K-major fp16, fp32 accumulate / output, optional fp16 copy of the output)
int gemm_f16_nt(const void* A16, int lda, const void* B16, int ldb, float* C, int ldc, void* C16, int ldc16, int M, int N,
                int K, const float* bias, const float* aux, int ldaux, float c_scale, int flags, cudaStream_t st,
                uint32_t* mask_out = nullptr, const uint32_t* mask_in = nullptr, int ldmw = 0, float* colsum = nullptr,
                float colsum_scale = 0.f,      // colsum[N] += colsum_scale * column sums of C16 (no bias allowed then)
                const F16Extra* ex = nullptr);
// C[M,N] += c_scale * A16[K,M]^T . B16[K,N]: fp16 operands as they lie in memory (M, N multiples of 64), split-K atomics
bool gemm_f16_tn_supported(int lda, int ldb, int M, int N);
// tn_R > 0 (multiple of 128, M = 2 tn_R): two-tap form, C rows [0, tn_R) = A16[k][m]^T . B16[k + tn_shift], rows [tn_R, 2 tn_R) =
// A16[k][m - tn_R]^T . B16[k]  (A16 has tn_R columns; B16 rows past K are zero-filled)
int gemm_f16_tn(const void* A16, int lda, const void* B16, int ldb, float* C, int ldc, int M, int N, int K, float c_scale,
                int split_k, cudaStream_t st, int tn_R = 0, int tn_shift = 0);
// out[i] = half(in[i])
int to_half(const float* in, void* out, int64_t n, cudaStream_t st);
int transpose_half(const float* in, int K, int N, void* out, int ldo, cudaStream_t st);
int transpose(const float* in, int ldi, float* out, int ldo, int rows, int cols, int round_out, cudaStream_t st);
int round_copy(const float* in, float* out, int64_t n, cudaStream_t st);

// img (nullable): pre-built weight image of the layer (block_images)
int block_fwd(const float* x, float* xout, float* zc, int ldz, const unsigned char* img,
              const float* wf, const float* wg, const float* dense, const float* prebias, const float* dense_bias,
              int M, int T, int d, int C, int is_last, cudaStream_t st);
int block_umma_set_trap_info(unsigned int* p);     // debug: see wn_debug_trap_info
int block_fwd_h_set_trap_info(unsigned int* p);
void set_fwd_h_timeline(long long* p);
// all forward layers in one persistent kernel (block_fwd_h.cu, "chain")
int block_fwd_chain(void* xs_ring, float* xall, float* zcat, void* zcat16, int ldz, const unsigned char* img,
                    const float* prebias, const float* dense_bias, const int* dilations, int L, int B, int T,
                    unsigned int* flags, cudaStream_t st, int ring = 0, int last_dense = 0, int zcols = 0);
// ring: slots of the split-row ring [ring][B][T][hi 32 | lo 32] (0: the default depth, forward only; L + 1: every layer's
// rows are kept -- what the backward chain reads); last_dense: the last layer also writes x' (xall needs L + 1 slots)
int64_t block_fwd_chain_ring_bytes(int64_t M, int ring = 0);
bool block_umma_enabled();
void set_block_timeline(long long* p);   // debug: clock64 stamps of block_fwd_umma CTA 0 (4 tiles x 8 phases)
void set_block_impl(int mma);
int block_fwd_umma(const float* x, float* xout, float* zc, int ldz, const unsigned char* img, const float* wf, const float* wg, const float* dense,
                   const float* prebias, const float* dense_bias, int B, int T, int d, int is_last, cudaStream_t st);
int block_bwd_pre_umma(const float* x, const float* dxn, const float* dZcat, const void* dZcat16, float dz_scale, int ldz,
                       int zcol, float* dpre, const unsigned char* img_pre, const float* prebias, int B, int T, int d,
                       int is_last, int pdl_next, cudaStream_t st);
// weight gradients of ALL layers in one persistent launch (after the pre / dx chain has finished)
int block_wgrad_all(const float* x, const float* dx, const float* dpre, const float* Zcat, int ldz, float* gwf, float* gwg,
                    float* gdense, float* gprebias, float* gdense_bias, const int* dilations, int L, int B, int T,
                    cudaStream_t st, int last_dense = 0);
int block_wgrad_umma(const float* x, const float* dxn, const float* dpre, const float* Zcat, int ldz, int zcol,
                     float* gwf, float* gwg, float* gdense, float* gprebias, float* gdense_bias, int B, int T, int d,
                     int is_last, int pdl, cudaStream_t st);
int block_bwd_dx_umma(const float* dxn, const float* dpre, float* dx, const unsigned char* img_dx, int B, int T, int d,
                      int is_last, int pdl_next, cudaStream_t st);
// tcgen05 GEMM when the shape allows, mma.sync otherwise (api.cu); mode as gemm_umma
int gemm_dispatch(int mode, GemmParams p, int split_k, cudaStream_t st);
// blocks of arbitrary channel widths built from GEMMs (block_generic.cu)
int64_t generic_scratch_floats(int64_t M, int R, int D);
int generic_block_fwd(const float* x, float* xout, float* zcat, int ldz, int zcol, float* P, const float* wf,
                      const float* wg, const float* dense, const float* prebias, const float* dense_bias,
                      float* scratch, int B, int T, int d, int R, int D, int is_last, cudaStream_t st);
int generic_block_bwd(const float* x, const float* dxn, const float* dZcat, const float* zcat, int ldz, int zcol,
                      const float* P, float* dx, float* dpre, const float* wf, const float* wg, const float* dense,
                      float* gwf, float* gwg, float* gdense, float* gprebias, float* gdense_bias, float* scratch,
                      int B, int T, int d, int R, int D, int is_last, cudaStream_t st);
// wide residual blocks (R, D multiples of 64) in 16-bit storage, built from the fp16 tcgen05 GEMM (block_wide16.cu)
bool wide16_supported(int R, int D);
int64_t wide16_images_bytes(int L, int R, int D);
int64_t wide16_wgrad_tmp_floats(int L, int R, int D);
int wide16_images(void* img, const float* filter, const float* gate, const float* dense, int L, int R, int D, cudaStream_t st);
int wide16_to_float(const void* in, float* out, float scale, int64_t n, cudaStream_t st);
int wide16_block_fwd(const void* x16, void* x16_out, void* P16, void* zcat16, int ldz, int zcol, const void* img_l,
                     const float* prebias, const float* dense_bias, int B, int T, int d, int R, int D, cudaStream_t st);
int wide16_block_bwd(const void* x16, const void* dxn16, const void* dzcat16, int ldz, int zcol, float cs, const void* P16,
                     const void* zcat16, void* dz16, void* dpre16, void* dx16_out, const void* img_l, float inv_scale,
                     float* wtmp, float* gdense, float* gprebias, float* gdense_bias_below, float* cs_scratch, int B, int T, int d,
                     int R, int D, cudaStream_t st);
int wide16_colsum_chunks();
int wide16_unpack_wgrad(const float* tmp, float* gwf, float* gwg, int L, int R, int D, cudaStream_t st);
// second-generation forward block (block_fwd_h.cu): fp16 split rows [hi 32 | lo 32] between layers
int64_t block_h_images_bytes(int L);
uint32_t block_h_img_stride();
int block_h_images(unsigned char* img, const float* filter, const float* gate, const float* dense, int L, cudaStream_t st);
int split_rows(const float* x, void* xs, int64_t M, cudaStream_t st);
int block_fwd_h(const void* xs_in, void* xs_out, float* xout, float* zcat, void* zcat16, int ldz, int zcol, const unsigned char* img,
                const float* prebias, const float* dense_bias, int B, int T, int d, int is_last, int pdl_next,
                cudaStream_t st);
// ---- backward chain on fp16 split rows (block_bwd_h.cu) ----
int64_t block_bwd_h_images_bytes(int L);
int block_bwd_h_images(unsigned char* img, const float* filter, const float* gate, const float* dense, int L, cudaStream_t st);
int64_t block_bwd_chain_flag_words(int L, int B, int T);
// pre-activation + input gradients of ALL layers in one persistent flag-ordered kernel (fp16 scaled domain)
int block_bwd_chain(const void* xs, void* dxs, void* p16, const void* dz16, int ldz, float cs, const unsigned char* img_f,
                    const unsigned char* img_b, const float* prebias, const int* dilations, int L, int B, int T,
                    unsigned int* flags, cudaStream_t st, int last_dense = 0);
// weight gradients of all layers from the fp16 tiles (x split rows, fp16 Zcat, dpre, dx split rows); scale: out of the scaled domain
// block_bwd_chain + block_wgrad_h_all as ONE launch (z is recomputed, nothing is re-read from HBM for the weight gradients)
int block_bwd_chain_fused(const void* xs, void* dxs, void* p16, const void* dz16, int ldz, float cs, const unsigned char* img_f,
                          const unsigned char* img_b, const float* prebias, const int* dilations, int L, int B, int T,
                          unsigned int* flags, float wscale, float* gwf, float* gwg, float* gdense, float* gprebias,
                          float* gdense_bias, void* scratch, cudaStream_t st, int last_dense = 0);
int64_t block_bwd_fused_scratch_bytes(int L, int B, int T);      // bytes of `scratch`
int block_wgrad_h_all(const void* xs, const void* dxs, const void* p16, const void* zcat16, int ldz, float scale, float* gwf,
                      float* gwg, float* gdense, float* gprebias, float* gdense_bias, const int* dilations, int L, int B,
                      int T, cudaStream_t st, int last_dense = 0);
int unsplit_rows(const void* xs, float* x, int64_t M, float scale, cudaStream_t st);
int block_bwd_h_set_trap_info(unsigned int* p);
void set_bwd_h_timeline(long long* p);
int64_t block_images_bytes(int L);
uint32_t block_img_off_pre();
uint32_t block_img_off_dx();
uint32_t block_img_stride();
int block_images(unsigned char* img, const float* filter, const float* gate, const float* dense, int L, cudaStream_t st);
int block_bwd(const float* x, const float* dxn, const float* dzs, int ldz, float* dx, float* dpre,
              const float* zc, const float* wf, const float* wg, const float* dense, const float* prebias,
              float* gwf, float* gwg, float* gdense, float* gprebias, float* gdense_bias, int M, int T,
              int d, int C, int is_last, cudaStream_t st);

int mulaw_encode(const float* audio, int64_t n, const float* thresholds, int Q, int32_t* ids, cudaStream_t st);
int mulaw_decode(const int32_t* ids, int64_t n, const float* lut, int Q, float* out, cudaStream_t st);

int frontend_fwd(const int32_t* ids, const float* wc, float* x0, int M, int T, int Q, int R, void* xs, cudaStream_t st);
int frontend_bwd(const int32_t* ids, const float* dx0, float* gwc, int M, int T, int Q, int R, cudaStream_t st);
// scalar_input front end, backward: gw[k][r] += sum_t x[t - lag_k] * dx0[t][r], lags as in causal_conv (width, dilation 1)
int scalar_frontend_bwd(const float* x, const float* dx0, float* gw, int M, int T, int R, int width, cudaStream_t st);

// g16 (optional): the gradient also as fp16 [M,Q], (softmax - onehot) * scale16
int softmax_xent(float* logits, const int32_t* ids, int M, int T, int Q, float scale, float* row_loss_partials,
                 int n_partials, float* loss_out, int write_grad, void* g16, float scale16, cudaStream_t st);

int xent_finalize(const float* partials, int n, float scale, float* loss_out, cudaStream_t st);
// postprocess2 GEMM + softmax cross entropy + fp16 gradient + bias-gradient column sums in one kernel (post_xent.cu);
// -2: shape not supported (needs Q == 256, K % 64 == 0)
int post2_xent(const void* A16, int lda, const void* W16, int ldw, const float* bias, const int32_t* ids, int M, int T, int K,
               int Q, float loss_scale, float* partials, float* loss_out, void* g16, float scale16, float* bias_grad,
               float colsum_scale, cudaStream_t st);

// prebias[l][b][2D] = [filter_bias_l | gate_bias_l] + emb[b] . [gc_filter_l | gc_gate_l]
int cond_bias_fwd(float* prebias, const float* filter_bias, const float* gate_bias, const float* gc_filter,
                  const float* gc_gate, const float* emb_table, const int32_t* gc_ids, int L, int B, int D,
                  int G, int card, cudaStream_t st);
int cond_bias_bwd(const float* gprebias, float* gfilter_bias, float* ggate_bias, const float* gc_filter,
                  const float* gc_gate, float* ggc_filter, float* ggc_gate, const float* emb_table,
                  float* gemb_table, const int32_t* gc_ids, int L, int B, int D, int G, int card,
                  cudaStream_t st);
int causal_conv(const float* x, const float* w, float* y, int M, int T, int cin, int cout, int width, int d,
                cudaStream_t st);
int skip_bias_sum(const float* skip_bias, int L, int S, float* out, cudaStream_t st);
// proba[q] = float(softmax(double(logits))[q])   (model.py:584-585: float64 softmax, cast back to float32), one row
int softmax_f64(const float* logits, int Q, float* proba, cudaStream_t st);
// *loss += coef * sum(params^2) / 2   (model.py:670-680: l2_regularization_strength * sum of tf.nn.l2_loss)
int add_l2(float* loss, const float* params, int64_t n, float coef, cudaStream_t st);
int bcast_rows(const float* src, int n, float* dst, int rows, cudaStream_t st);
int add_inplace(float* dst, const float* src, int64_t n, int round_out, cudaStream_t st);
int relu_mask_add(float* dst, const float* grad, const float* act, int64_t n, int round_out, cudaStream_t st);

int optim_adam(float* w, const float* g, float* m, float* v, int64_t n, double lr_t, double beta1, double beta2,
               double eps, float l2, float gscale, cudaStream_t st);
int optim_momentum(float* w, const float* g, float* a, int64_t n, double lr, double mu, float l2, float gscale,
                   cudaStream_t st);
int optim_rmsprop(float* w, const float* g, float* ms, float* mom, int64_t n, double lr, double decay, double mu,
                  double eps, float l2, float gscale, cudaStream_t st);

}  // namespace wn
