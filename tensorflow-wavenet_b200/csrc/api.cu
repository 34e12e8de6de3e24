// extern "C" surface of libwavenet_b200.so: argument checking, the flat parameter layout,
// workspace carving and the launch sequences of the training / forward graphs.
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/wavenet_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "umma_common.cuh"
#include "gen_common.h"

namespace wn {

// ---------------------------------------------------------------------------------------
// per-kernel timing: one CUDA event after every launch of the training sequence
// ---------------------------------------------------------------------------------------
namespace {
constexpr int PROF_MAX = 4096;
bool g_prof_on = false;
int g_prof_n = 0;
cudaEvent_t g_prof_ev[PROF_MAX];
int g_prof_tag[PROF_MAX];
bool g_prof_created = false;
}  // namespace

void prof_mark(cudaStream_t st, int tag) {
  static int dbg = -1;
  if (dbg < 0) dbg = getenv("WN_DEBUG_SYNC") ? 1 : 0;
  if (dbg) {   // debugging aid: synchronise after every launch and report the first failing kernel kind
    cudaError_t e = cudaStreamSynchronize(st);
    fprintf(stderr, "[wn debug] after tag %d: %s\n", tag, cudaGetErrorString(e));
  }
  static int sync_tag = -2;     // WN_SYNC_TAG=<ProfTag>: synchronise after every launch of that kind (race hunting)
  if (sync_tag == -2) { const char* e = getenv("WN_SYNC_TAG"); sync_tag = e ? atoi(e) : -1; }
  if (sync_tag == tag) cudaStreamSynchronize(st);
  if (!g_prof_on || g_prof_n >= PROF_MAX) return;
  // inside a stream capture the record becomes an external event node, so that the graph replay stamps a
  // real, timeable event after every kernel node (an eager pass is host-launch-bound for the short kernels)
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(g_prof_ev[g_prof_n], st, cudaEventRecordExternal);
  else cudaEventRecord(g_prof_ev[g_prof_n], st);
  g_prof_tag[g_prof_n] = tag;
  ++g_prof_n;
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
// wn_set_grad_ready_event: recorded when the gradients of the tail bucket [skip, total) are final (see make_layout)
static cudaEvent_t g_tail_ready_event = nullptr;
static inline bool fused_blocks(const wn_config* c) {
  return c->residual_channels == c->dilation_channels && (c->residual_channels == 16 || c->residual_channels == 32);
}

static int check_cfg(const wn_config* c) {
  if (!c) return -1;
  if (c->n_layers < 1 || c->n_layers > WN_MAX_LAYERS) return -1;
  // R = D in {16, 32}: fused block kernels; any other widths (multiples of 4): GEMM-built blocks (block_generic.cu)
  if (c->residual_channels < 4 || (c->residual_channels & 3) || c->residual_channels > 256 ||
      (256 % c->residual_channels)) return -2;
  if (c->dilation_channels < 4 || (c->dilation_channels & 3) || c->dilation_channels > 1024) return -2;
  if (c->skip_channels < 4 || (c->skip_channels & 3)) return -2;
  if (c->quantization_channels < 4 || (c->quantization_channels & 3) || c->quantization_channels > 1024) return -2;
  if (c->gc_channels < 0 || c->gc_channels > 1024) return -2;
  if (c->gc_channels > 0 && c->gc_cardinality < 0) return -2;
  if (c->scalar_input && (c->initial_filter_width < 1 || c->initial_filter_width > 1024)) return -2;
  for (int i = 0; i < c->n_layers; ++i)
    if (c->dilations[i] < 1) return -1;
  return 0;
}

static int make_layout(const wn_config* c, wn_layout* o) {
  int rc = check_cfg(c);
  if (rc) return rc;
  const int64_t L = c->n_layers, R = c->residual_channels, D = c->dilation_channels, S = c->skip_channels,
                Q = c->quantization_channels, G = c->gc_channels;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t r = off; off = align_up(off + n, 64); return r; };
  // Group order = all-reduce buckets (wavenet/train_step.py): everything the residual-block backward produces comes first;
  // the skip / post-processing weights and biases -- 80 % of the bytes, final more than a millisecond before the end of
  // the step (their gradient GEMMs run first) -- form the contiguous tail [skip, total), which is reduced while the
  // backward chain still runs.
  o->causal = take(c->scalar_input ? (int64_t)c->initial_filter_width * R : 2 * Q * R);
  o->filter = take(L * 2 * R * D);
  o->gate = take(L * 2 * R * D);
  o->dense = take(L * D * R);
  o->gc_filter = G > 0 ? take(L * G * D) : -1;
  o->gc_gate = G > 0 ? take(L * G * D) : -1;
  o->filter_bias = c->use_biases ? take(L * D) : -1;
  o->gate_bias = c->use_biases ? take(L * D) : -1;
  o->dense_bias = c->use_biases ? take(L * R) : -1;
  o->gc_embedding = (G > 0 && c->gc_cardinality > 0) ? take((int64_t)c->gc_cardinality * G) : -1;
  o->skip = take(L * D * S);
  o->skip_bias = c->use_biases ? take(L * S) : -1;
  o->post1 = take(S * S);
  o->post2 = take(S * Q);
  o->post1_bias = c->use_biases ? take(S) : -1;
  o->post2_bias = c->use_biases ? take(Q) : -1;
  o->total = off;
  return 0;
}

struct Workspace {
  int32_t* ids;
  float* X;        // training: L buffers [M,R]; forward: 2 buffers
  float* Zcat;     // [M, L*D]
  float* A1;       // [M,S] tf32(relu(skip sum))
  float* A2;       // [M,S] tf32(relu(post1))
  float* S0;       // [M,S] raw skip sum            (residual_postproc only)
  float* T2;       // [M,S] tf32(A2 + S0)           (residual_postproc only)
  float* logits;   // [M,Q]                         (training only; forward writes to caller's)
  float* G1;       // [M,S]
  float* G2;       // [M,S]
  float* G3;       // [M,S]                         (residual_postproc only)
  float* dZcat;    // [M, L*D]   (tf32 gradient chain)
  void* dZcat16;   // [M, L*D] fp16, scaled by gscale (fp16 gradient chain; then dZcat, G1, G2 are not allocated)
  float* dX;       // tcgen05 path: L x [M,R], one per layer (the side-stream weight-gradient kernels read them
                   // long after the chain has moved on; no buffer is ever reused within a step); else 2 x [M,R]
  float* dpre;     // tcgen05 path: L x [M,2D]; else [M,2D]
  float* prebias;  // [L,B,2D]
  float* gprebias; // [L,B,2D]
  float* bsum;     // [S]
  float* gtmp;     // [S]
  float* partials; // [4096]
  float* WskipR;   // [L*D, S]  tf32-rounded weight copies (tcgen05 truncates raw fp32 operands)
  float* W1R;      // [S, S]
  float* W2R;      // [S, Q]
  unsigned char* Wimg;  // per-layer weight images of the tcgen05 block kernels (C == 32)
  unsigned char* WimgH; // per-layer fp16 split weight images of the forward block (block_fwd_h.cu)
  void *Zcat16, *A1h, *A2h;       // fp16 copies of the forward GEMM A operands (fp16 forward chain)
  void *Wskip16, *W1h, *W2h;      // fp16 K-major weight copies: [S][L*D], [S][S], [Q][S]
  void *dlog16, *G1h, *G2h;       // gradients of the post-processing chain as fp16, scaled by gscale (training, fp16 chain)
  float* cs_scratch;              // per-chunk partial column sums (bias gradients of the fp16 chain)
  uint32_t *maskA1, *maskA2;      // relu masks of the skip sum / postprocess1 outputs as bits, [M][S/32]
  void *Wskipg, *W1g, *W2g;       // fp16 copies of the weights as stored ([L*D][S], [S][S], [S][Q]): input-gradient operands
  unsigned int* chain_flags;      // [L][B * ceil(T/128)] tile flags of the persistent forward kernel (null: per-layer launches)
  void* XS;             // 2 x [M][hi 32 | lo 32] fp16 split rows: the residual stream between forward layers
  int umma_bwd;    // 1 when the tcgen05 backward path is used
  int bwd16;       // 1: fp16 backward chain (block_bwd_h.cu): XS keeps every layer's split rows, DXS / P16 hold dx / dpre
  void *DXS, *P16;                // [L][M][hi 32 | lo 32] input gradients, [L][M][df 32 | dg 32] pre-activation gradients (fp16, scaled)
  unsigned char* WimgB;           // per-layer weight images of the backward chain
  unsigned int* bflags;           // tile flags + work counter of the backward chain
  void* bpart;                    // per-CTA weight-gradient partial sums of the fused backward chain
  int wide16;      // 1: wide residual blocks in 16-bit storage (block_wide16.cu: R, D multiples of 64)
  void *X16, *P16w, *Wimg16;      // [L or 2][M][R] layer inputs, [L or 1][M][2D] pre-activations (fp16), per-layer weight images
  void *dz16w, *dpre16w, *dx16w;  // backward temporaries: [M][D], [M][2D], 2 x [M][R] (fp16, scaled domain)
  float *wtmp16, *cs_scratch2;    // [L][2][R][2D] filter | gate gradients before unpacking; column-sum partials
  float* Pall;     // generic-width blocks: saved pre-activations [L or 1][M][2D]
  float* gscratch; // generic-width blocks: operand splits / temporaries (generic_scratch_floats)
  int64_t bytes;
};

static bool fwd_h_enabled();
// WN_NO_SIDE=1 keeps every kernel of the step on the caller's stream (debugging aid)
static int side_mode() {      // WN_NO_SIDE: 1 = no side streams, 2 = no post-processing side stream, 3 = no block side stream
  static int v = -1;
  if (v < 0) { const char* e = getenv("WN_NO_SIDE"); v = e ? atoi(e) : 0; }
  return v;
}
static bool no_side_streams() { return side_mode() == 1; }
static bool bwd_pdl_enabled() {      // WN_BWD_PDL=0: plain launches in the backward chain
  static int v = -1;
  if (v < 0) { const char* e = getenv("WN_BWD_PDL"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v == 1;
}
// forward GEMM chain (skip sum, postprocess1, postprocess2) on fp16 operands (default) or tf32 (WN_FWD_GEMM=tf32)
static bool fwd16_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WN_FWD_GEMM");
    const char* g = getenv("WN_GEMM_IMPL");
    v = ((e && strcmp(e, "tf32") == 0) || (g && strcmp(g, "mma") == 0)) ? 0 : 1;
  }
  return v == 1;
}
// forward layers: one persistent kernel ordered by tile flags (default) or one launch per layer (WN_FWD_CHAIN=0)
static bool fwd_chain_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WN_FWD_CHAIN");
    v = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v == 1;
}
// backward of the fused blocks: one persistent fp16 chain kernel + one weight-gradient launch (default), or the
// first-generation per-layer TF32 kernels (WN_BWD_CHAIN=0)
static bool fused_xent_enabled() {      // WN_FUSED_XENT=0: postprocess2 GEMM, cross-entropy kernel and column sums as separate launches
  static const bool on = [] { const char* e = getenv("WN_FUSED_XENT"); return !(e && e[0] == '0'); }();
  return on;
}
static bool bwd_fused_enabled() {      // WN_BWD_FUSED=0: weight gradients of the residual blocks as their own launch
  static const bool on = [] { const char* e = getenv("WN_BWD_FUSED"); return !(e && e[0] == '0'); }();
  return on;
}
static bool wide16_enabled() {      // WN_WIDE16=0: wide blocks through the fp32 / TF32 GEMM-built path (block_generic.cu)
  static const bool on = [] { const char* e = getenv("WN_WIDE16"); return !(e && e[0] == '0'); }();
  return on;
}
static bool bwd_chain_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WN_BWD_CHAIN");
    v = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v == 1;
}
static void carve(const wn_config* c, int B, int T, bool training, void* base, Workspace* w) {
  const int64_t M = (int64_t)B * T, L = c->n_layers, R = c->residual_channels, D = c->dilation_channels,
                S = c->skip_channels, Q = c->quantization_channels;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t r = off;
    off = align_up(off + bytes, 256);
    return base ? (void*)((char*)base + r) : (void*)nullptr;
  };
  const int64_t f = sizeof(float);
  const bool umma_blocks = fused_blocks(c) && (R == 32) && block_umma_enabled();
  const bool fwd_h = umma_blocks && fwd_h_enabled();
  const bool chain = fwd_h && fwd_chain_enabled();
  const bool wide16 = !fused_blocks(c) && wide16_supported((int)R, (int)D) && wide16_enabled() && fwd16_enabled() &&
                      !c->residual_postproc && (S % 8) == 0 && (!training || ((Q % 64) == 0 && (S % 64) == 0));
  w->wide16 = wide16 ? 1 : 0;
  const bool f16_chain = (fwd_h || wide16) && fwd16_enabled() && !c->residual_postproc && ((L * D) % 8) == 0 && (S % 8) == 0;
  // fp16 gradient chain of the post-processing layers: all-or-nothing (its GEMMs keep no fp32 copies of G1 / G2), so
  // every weight-gradient shape must suit the fp16 MN-major form (multiples of 64); otherwise the tf32 chain runs
  const bool g16 = training && f16_chain && (Q % 64) == 0 && (S % 64) == 0 && ((L * D) % 64) == 0;
  const bool bwd16 = g16 && chain && bwd_chain_enabled();
  w->bwd16 = bwd16 ? 1 : 0;
  w->ids = (int32_t*)take(M * 4 + 16);
  w->X = (float*)take((training && !bwd16 && !wide16 ? L : 2) * M * R * f);
  w->Zcat = (float*)take((bwd16 || wide16) ? 256 : M * L * D * f);      // (fp16 backward chain: every consumer of z reads Zcat16)
  w->A1 = (float*)take(M * S * f);
  w->A2 = (float*)take(M * S * f);
  w->S0 = c->residual_postproc ? (float*)take(M * S * f) : nullptr;
  w->T2 = c->residual_postproc ? (float*)take(M * S * f) : nullptr;
  w->prebias = (float*)take(L * B * 2 * D * f);
  w->bsum = (float*)take(S * f);
  w->partials = (float*)take(4096 * f);
  w->WskipR = (float*)take(S * L * D * f);
  w->W1R = (float*)take(S * S * f);
  w->W2R = (float*)take(Q * S * f);
  w->X16 = w->P16w = w->Wimg16 = w->dz16w = w->dpre16w = w->dx16w = nullptr;
  w->wtmp16 = w->cs_scratch2 = nullptr;
  if (wide16) {
    w->X16 = take((training ? L : 2) * M * R * 2);
    w->P16w = take((training ? L : 1) * M * 2 * D * 2);
    w->Wimg16 = take(wide16_images_bytes((int)L, (int)R, (int)D));
    if (training) {
      w->dz16w = take(M * D * 2);
      w->dpre16w = take(M * 2 * D * 2);
      w->dx16w = take(2 * M * R * 2);
      w->wtmp16 = (float*)take(wide16_wgrad_tmp_floats((int)L, (int)R, (int)D) * f);
      w->cs_scratch2 = (float*)take((int64_t)B * wide16_colsum_chunks() * 2 * D * f);
    }
  }
  if (!fused_blocks(c) && !wide16) {
    w->Pall = (float*)take((training ? L : 1) * M * 2 * D * f);
    w->gscratch = (float*)take(generic_scratch_floats(M, (int)R, (int)D) * f);
  } else {
    w->Pall = w->gscratch = nullptr;
  }
  w->Wimg = umma_blocks ? (unsigned char*)take(block_images_bytes((int)L)) : nullptr;
  w->umma_bwd = (training && umma_blocks) ? 1 : 0;
  w->WimgH = fwd_h ? (unsigned char*)take(block_h_images_bytes((int)L)) : nullptr;
  w->XS = fwd_h ? take(chain ? block_fwd_chain_ring_bytes(M, bwd16 ? (int)(L > 2 ? L : 2) : 0) : 2 * M * 128) : nullptr;
  w->chain_flags = chain ? (unsigned int*)take((L * (int64_t)B * ((T + 127) / 128) + 1) * 4) : nullptr;
  if (bwd16) {
    w->DXS = take(L * M * 128);
    w->P16 = take(L * M * 128);
    w->WimgB = (unsigned char*)take(block_bwd_h_images_bytes((int)L));
    w->bflags = (unsigned int*)take(block_bwd_chain_flag_words((int)L, B, T) * 4);
    w->bpart = bwd_fused_enabled() ? take(block_bwd_fused_scratch_bytes((int)L, B, T)) : nullptr;
  } else {
    w->DXS = w->P16 = nullptr;
    w->WimgB = nullptr;
    w->bflags = nullptr;
    w->bpart = nullptr;
  }
  if (f16_chain) {
    w->Zcat16 = take(M * L * D * 2);
    w->A1h = take(M * S * 2);
    w->A2h = take(M * S * 2);
    w->Wskip16 = take(S * L * D * 2);
    w->W1h = take(S * S * 2);
    w->W2h = take(Q * S * 2);
  } else {
    w->Zcat16 = w->A1h = w->A2h = w->Wskip16 = w->W1h = w->W2h = nullptr;
  }
  if (g16) {      // (not a pointer test: a size query carves from a null base)
    w->dlog16 = take(M * Q * 2);
    w->dZcat16 = take(M * L * D * 2);
    w->G1h = take(M * S * 2);
    w->G2h = take(M * S * 2);
    w->Wskipg = take(L * D * S * 2);
    w->W1g = take(S * S * 2);
    w->W2g = take(S * Q * 2);
    w->maskA1 = (uint32_t*)take(M * (S / 32) * 4);
    w->maskA2 = (uint32_t*)take(M * (S / 32) * 4);
    w->cs_scratch = (float*)take(colsum16_scratch_floats((int)(S > Q ? S : Q)) * f);
  } else {
    w->cs_scratch = nullptr;
    w->dlog16 = w->G1h = w->G2h = w->Wskipg = w->W1g = w->W2g = w->dZcat16 = nullptr;
    w->maskA1 = w->maskA2 = nullptr;
  }
  if (training) {
    w->logits = (float*)take(M * Q * f);
    w->G1 = g16 ? nullptr : (float*)take(M * S * f);
    w->G2 = g16 ? nullptr : (float*)take(M * S * f);
    w->G3 = c->residual_postproc ? (float*)take(M * S * f) : nullptr;
    w->dZcat = g16 ? nullptr : (float*)take(M * L * D * f);
    w->dX = (float*)take((bwd16 ? 1 : w->umma_bwd ? L : 2) * M * R * f);
    w->dpre = (float*)take((bwd16 || wide16) ? 256 : (w->umma_bwd ? L : 1) * M * 2 * D * f);
    w->gprebias = (float*)take(L * B * 2 * D * f);
    w->gtmp = (float*)take(S * f);
  } else {
    w->logits = w->G1 = w->G2 = w->G3 = w->dZcat = w->dX = w->dpre = w->gprebias = w->gtmp = nullptr;
  }
  w->bytes = off;
}

// forward block implementation: fp16 split rows (default) or the TF32 3-term kernel (WN_BLOCK_FWD=tf32)
static bool fwd_h_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WN_BLOCK_FWD");
    v = (e && strcmp(e, "tf32") == 0) ? 0 : 1;
  }
  return v == 1;
}

static inline const float* P(const float* base, int64_t off) { return off >= 0 ? base + off : nullptr; }
static inline float* P(float* base, int64_t off) { return off >= 0 ? base + off : nullptr; }

static GemmParams gp(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  return p;
}

static int split_for(int m_out, int n_out, int k) {
  const int tiles = ((m_out + 127) / 128) * ((n_out + 127) / 128);
  const int nkb = (k + 31) / 32;
  int s = (4 * sm_count() + tiles - 1) / tiles;
  if (s > nkb / 4) s = nkb / 4;
  if (s < 1) s = 1;
  if (s > 65535) s = 65535;
  return s;
}

// mode 0 NN: C = A[M,K].B[K,N]   1 NT: C = A[M,K].B[N,K]^T   2 TN: C += A[K,M]^T.B[K,N] (split-K, atomics).
// tcgen05 by default; WN_GEMM_IMPL=mma (or a shape the tcgen05 kernel does not take) selects the mma.sync
// kernel -- also how one implementation is validated against the other.
static int g_gemm_impl = -1;   // -1 unset, 0 tcgen05, 1 mma.sync
static bool use_mma_gemm() {
  if (g_gemm_impl < 0) {
    const char* e = getenv("WN_GEMM_IMPL");
    g_gemm_impl = (e && strcmp(e, "mma") == 0) ? 1 : 0;
  }
  return g_gemm_impl == 1;
}
int gemm_dispatch(int mode, GemmParams p, int split_k, cudaStream_t st);
static int gemm(int mode, GemmParams p, int split_k, cudaStream_t st) { return gemm_dispatch(mode, p, split_k, st); }
int gemm_dispatch(int mode, GemmParams p, int split_k, cudaStream_t st) {
  if (mode == 2) p.flags |= GEMM_ATOMIC;
  if (!use_mma_gemm() && gemm_umma_supported(mode, p)) return gemm_umma(mode, p, nullptr, 0, split_k, st);
  return gemm_tf32(mode, p, split_k, st);
}

#define RC(x)            \
  do {                   \
    int rc__ = (x);      \
    if (rc__) return rc__; \
  } while (0)

// network forward from ids: X (all layers when training), Zcat, A1, A2 and logits
// debug (tools/stage_times.py): WN_TRUNCATE=k cuts the step's launch sequence after stage k -- 1 forward layers,
// 2 forward GEMMs + loss, 3 post-processing gradient GEMMs, 4 layer backward -- so that graph replay times of the
// prefixes give the real incremental cost of every stage (event-node kernel times undo the overlap between launches)
static int truncate_stage() {
  const char* e = getenv("WN_TRUNCATE");
  return e ? atoi(e) : 0;
}

static int run_forward(const wn_config* c, const wn_layout& lo, const float* params, Workspace& w,
                       const int32_t* ids, const int32_t* gc_ids, int B, int T, bool training, float* logits,
                       cudaStream_t st, const float* scalar_in = nullptr, bool last_only = false) {
  const int M = B * T, L = c->n_layers, R = c->residual_channels, D = c->dilation_channels,
            S = c->skip_channels, Q = c->quantization_channels, G = c->gc_channels;
  const int ldz = L * D;
  if (G > 0 && !gc_ids) return -1;
  // Everything that only converts parameters (weight images of the backward kernels, fp16 / transposed weight copies,
  // the skip bias sum: ~10 launches of a few microseconds each) runs on a side branch next to the forward-layer kernel
  // -- its CTAs fit beside the two resident chain CTAs of an SM -- and is joined in front of the GEMMs.
  static cudaStream_t prep = nullptr;
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  const bool use_prep = training && w.chain_flags && w.Zcat16 && !g_prof_on && !no_side_streams();
  if (use_prep && !prep) {
    RC((int)cudaStreamCreateWithFlags(&prep, cudaStreamNonBlocking));
    RC((int)cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    RC((int)cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  cudaStream_t ps = use_prep ? prep : st;
  const float* bsum = nullptr;
  auto prep_weights = [&]() -> int {
    if (c->use_biases) {
      RC(skip_bias_sum(params + lo.skip_bias, L, S, w.bsum, ps));
      prof_mark(ps, PT_SKIP_BIAS_SUM);
      bsum = w.bsum;
    }
    // tf32-rounded weight copies: B operands of the TF32 forward products and of the TF32 input-gradient products
    // (nothing reads them when both directions run the fp16 chain)
    if (!w.Zcat16 || (training && !w.dlog16)) {
      RC(round_copy(params + lo.skip, w.WskipR, (int64_t)ldz * S, ps));
      RC(round_copy(params + lo.post1, w.W1R, (int64_t)S * S, ps));
      RC(round_copy(params + lo.post2, w.W2R, (int64_t)S * Q, ps));
    }
    if (w.Zcat16) {   // fp16 forward chain: same 11-bit operand mantissas as tf32, half the L2 -> SM operand bytes
      RC(transpose_half(params + lo.skip, ldz, S, w.Wskip16, ldz, ps));
      RC(transpose_half(params + lo.post1, S, S, w.W1h, S, ps));
      RC(transpose_half(params + lo.post2, S, Q, w.W2h, S, ps));
      if (w.dlog16) {   // operands of the fp16 input-gradient chain (the weights as stored)
        RC(to_half(params + lo.post2, w.W2g, (int64_t)S * Q, ps));
        RC(to_half(params + lo.post1, w.W1g, (int64_t)S * S, ps));
        RC(to_half(params + lo.skip, w.Wskipg, (int64_t)ldz * S, ps));
      }
    }
    prof_mark(ps, PT_MISC);
    return 0;
  };
  if (use_prep) {
    RC((int)cudaEventRecord(ev_fork, st));
    RC((int)cudaStreamWaitEvent(prep, ev_fork, 0));
    if (w.bwd16) RC(block_bwd_h_images(w.WimgB, params + lo.filter, params + lo.gate, params + lo.dense, L, prep));
    else RC(block_images(w.Wimg, params + lo.filter, params + lo.gate, params + lo.dense, L, prep));      // backward images
    RC(prep_weights());
    RC((int)cudaEventRecord(ev_join, prep));
    RC(block_h_images(w.WimgH, params + lo.filter, params + lo.gate, params + lo.dense, L, st));
  } else if (w.Wimg) {   // first, so that at least two launches separate it from the first block kernel (PDL, common.cuh)
    if (training && w.bwd16) RC(block_bwd_h_images(w.WimgB, params + lo.filter, params + lo.gate, params + lo.dense, L, st));
    else RC(block_images(w.Wimg, params + lo.filter, params + lo.gate, params + lo.dense, L, st));
    if (w.WimgH) RC(block_h_images(w.WimgH, params + lo.filter, params + lo.gate, params + lo.dense, L, st));
    prof_mark(st, PT_MISC);
  }
  RC(cond_bias_fwd(w.prebias, P(params, lo.filter_bias), P(params, lo.gate_bias), P(params, lo.gc_filter),
                   P(params, lo.gc_gate), P(params, lo.gc_embedding), gc_ids, L, B, D, gc_ids ? G : 0, c->gc_cardinality, st));
  prof_mark(st, PT_COND_BIAS);
  // (residual stream as fp16 split rows between the forward layers: the front end writes both forms of its output)
  if (c->scalar_input) {      // model.py:143-153: a width-IFW causal convolution of the raw waveform (1 input channel)
    if (!scalar_in) return -1;
    RC(causal_conv(scalar_in, params + lo.causal, w.X, M, T, 1, R, c->initial_filter_width, 1, st));
    if (w.WimgH) RC(split_rows(w.X, w.XS, M, st));
  } else {
    RC(frontend_fwd(ids, params + lo.causal, w.X, M, T, Q, R, w.WimgH ? w.XS : nullptr, st));
  }
  prof_mark(st, PT_FRONTEND_FWD);
  const int64_t xs = (int64_t)M * R;
  if (w.wide16) {      // 16-bit storage between the layers; weight images of every layer (forward and backward forms)
    RC(wide16_images(w.Wimg16, params + lo.filter, params + lo.gate, params + lo.dense, L, R, D, st));
    RC(to_half(w.X, w.X16, xs, st));
    prof_mark(st, PT_MISC);
  }
  if (w.chain_flags) {
    // fp16 backward chain: the split rows of EVERY layer stay (ring = L) and nothing reads fp32 x' / z
    const bool b16 = training && w.bwd16;
    RC(block_fwd_chain(w.XS, (training && !b16) ? w.X : nullptr, b16 ? nullptr : w.Zcat, w.Zcat16, ldz, w.WimgH, w.prebias,
                       lo.dense_bias >= 0 ? params + lo.dense_bias : nullptr, c->dilations, L, B, T, w.chain_flags, st,
                       b16 ? (L > 2 ? L : 2) : 0));
  }
  for (int l = 0; l < L && !w.chain_flags; ++l) {
    const float* xin = training ? w.X + l * xs : w.X + (l & 1) * xs;
    float* xout = training ? w.X + (l + 1) * xs : w.X + ((l + 1) & 1) * xs;
    const int last = (l == L - 1);
    if (w.WimgH) {
      char* xs_in = (char*)w.XS + (int64_t)(l & 1) * M * 128;
      char* xs_out = (char*)w.XS + (int64_t)((l + 1) & 1) * M * 128;
      RC(block_fwd_h(xs_in, last ? nullptr : xs_out, last ? nullptr : xout, w.Zcat, w.Zcat16, ldz, l * D,
                     w.WimgH + (size_t)l * block_h_img_stride(), w.prebias + (int64_t)l * B * 2 * D,
                     lo.dense_bias >= 0 ? params + lo.dense_bias + (int64_t)l * R : nullptr, B, T, c->dilations[l], last,
                     /*pdl_next=*/!last, st));
      continue;
    }
    if (w.wide16) {
      char* x16_in = (char*)w.X16 + (training ? (int64_t)l : (int64_t)(l & 1)) * xs * 2;
      char* x16_out = (char*)w.X16 + (training ? (int64_t)(l + 1) : (int64_t)((l + 1) & 1)) * xs * 2;
      RC(wide16_block_fwd(x16_in, last ? nullptr : x16_out, (char*)w.P16w + (training ? (int64_t)l : 0) * M * 2 * D * 2, w.Zcat16, ldz,
                          l * D, (char*)w.Wimg16 + (int64_t)l * (wide16_images_bytes(1, R, D)), w.prebias + (int64_t)l * B * 2 * D,
                          lo.dense_bias >= 0 ? params + lo.dense_bias + (int64_t)l * R : nullptr, B, T, c->dilations[l], R, D, st));
      continue;
    }
    if (w.gscratch) {
      RC(generic_block_fwd(xin, last ? nullptr : xout, w.Zcat, ldz, l * D, w.Pall + (training ? (int64_t)l * M * 2 * D : 0),
                           params + lo.filter + (int64_t)l * 2 * R * D, params + lo.gate + (int64_t)l * 2 * R * D,
                           params + lo.dense + (int64_t)l * D * R, w.prebias + (int64_t)l * B * 2 * D,
                           lo.dense_bias >= 0 ? params + lo.dense_bias + (int64_t)l * R : nullptr, w.gscratch, B, T,
                           c->dilations[l], R, D, last, st));
      continue;
    }
    RC(block_fwd(xin, last ? nullptr : xout, w.Zcat + (int64_t)l * D, ldz,
                 w.Wimg ? w.Wimg + (size_t)l * block_img_stride() : nullptr, params + lo.filter + (int64_t)l * 2 * R * D,
                 params + lo.gate + (int64_t)l * 2 * R * D, params + lo.dense + (int64_t)l * D * R,
                 w.prebias + (int64_t)l * B * 2 * D, lo.dense_bias >= 0 ? params + lo.dense_bias + (int64_t)l * R : nullptr,
                 M, T, c->dilations[l], R, last, st));
  }
  if (use_prep) RC((int)cudaStreamWaitEvent(st, ev_join, 0));      // join the parameter-conversion branch
  else RC(prep_weights());
  if (training && truncate_stage() == 1) return 0;
  // predict_proba (model.py:564-590) needs the LAST row of the logits only: the post-processing runs on that one row
  const int Mg = last_only ? 1 : M;
  const int64_t r0 = last_only ? (int64_t)M - 1 : 0;
  if (w.Zcat16) {
    // fp32 copies of the two hidden activations only where something reads them: the tf32 gradient chain
    float* a1_32 = (training && !w.dlog16) ? w.A1 : nullptr;
    float* a2_32 = (training && !w.dlog16) ? w.A2 : nullptr;
    const char* z16 = (const char*)w.Zcat16 + r0 * ldz * 2;
    RC(gemm_f16_nt(z16, ldz, w.Wskip16, ldz, a1_32, S, w.A1h, S, Mg, S, ldz, bsum, nullptr, 0, 1.f, GEMM_RELU | GEMM_ROUND, st,
                   training ? w.maskA1 : nullptr, nullptr, S / 32));
    prof_mark(st, PT_GEMM_SKIP_FWD);
    RC(gemm_f16_nt(w.A1h, S, w.W1h, S, a2_32, S, w.A2h, S, Mg, S, S, P(params, lo.post1_bias), nullptr, 0, 1.f, GEMM_RELU | GEMM_ROUND, st,
                   training ? w.maskA2 : nullptr, nullptr, S / 32));
    prof_mark(st, PT_GEMM_POST1_FWD);
    if (!logits) return 0;      // training with the fused postprocess2 + cross-entropy kernel: the caller runs it
    RC(gemm_f16_nt(w.A2h, S, w.W2h, S, logits, Q, nullptr, 0, Mg, Q, S, P(params, lo.post2_bias), nullptr, 0, 1.f, 0, st));
    prof_mark(st, PT_GEMM_POST2_FWD);
    return 0;
  }
  {  // total = sum_l skip_l  ->  relu            (model.py:430-431)
    GemmParams p = gp(w.Zcat + r0 * ldz, ldz, w.WskipR, S, w.A1, S, Mg, S, ldz);
    p.bias = bsum;
    p.flags = GEMM_RELU | GEMM_ROUND;
    if (c->residual_postproc) { p.C2 = w.S0; p.ldc2 = S; }
    RC(gemm(0, p, 1, st));
    prof_mark(st, PT_GEMM_SKIP_FWD);
  }
  {  // conv1 -> relu                             (model.py:432-435)
    GemmParams p = gp(w.A1, S, w.W1R, S, w.A2, S, Mg, S, S);
    p.bias = P(params, lo.post1_bias);
    p.flags = GEMM_RELU | GEMM_ROUND;
    RC(gemm(0, p, 1, st));
    prof_mark(st, PT_GEMM_POST1_FWD);
  }
  const float* x2 = w.A2;
  if (c->residual_postproc) {  // transformed2 += total   (model.py:436-437)
    RC((int)cudaMemcpyAsync(w.T2, w.A2, (size_t)Mg * S * sizeof(float), cudaMemcpyDeviceToDevice, st));
    RC(add_inplace(w.T2, w.S0, (int64_t)Mg * S, 1, st));
    x2 = w.T2;
  }
  {  // conv2                                      (model.py:438-440)
    GemmParams p = gp(x2, S, w.W2R, Q, logits, Q, Mg, Q, S);
    p.bias = P(params, lo.post2_bias);
    RC(gemm(0, p, 1, st));
    prof_mark(st, PT_GEMM_POST2_FWD);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// Stand-alone residual block entry points (wn_block_fwd / wn_block_bwd) on the PRODUCTION kernels: the same
// block_fwd_chain_kernel / block_bwd_pre_umma<fp16 dz> / block_bwd_dx_umma / block_wgrad_all launches that
// wn_loss_grad issues, driven as a one-layer network.  Scratch (split rows, weight images, tile flags, fp16 copy of the
// skip-path gradient) comes from the stream-ordered allocator: these two entry points are test / integration surfaces,
// the training step itself never allocates.
// ---------------------------------------------------------------------------------------
struct StreamScratch {
  cudaStream_t st; char* base = nullptr; int64_t off = 0, cap = 0;
  int init(int64_t bytes, cudaStream_t s) { st = s; cap = bytes; return (int)cudaMallocAsync((void**)&base, (size_t)bytes, s); }
  void* take(int64_t bytes) { void* p = base + off; off = align_up(off + bytes, 1024); return off <= cap ? p : nullptr; }
  ~StreamScratch() { if (base) cudaFreeAsync(base, st); }
};

static int block_fwd_production(const float* x, float* x_out, float* zcat, int ldz, const float* filter, const float* gate,
                                const float* dense, const float* prebias, const float* dense_bias, int B, int T, int d,
                                int is_last, cudaStream_t st) {
  const int64_t M = (int64_t)B * T;
  const int64_t n_tiles = (int64_t)B * ((T + 127) / 128);
  StreamScratch sc;
  RC(sc.init(2 * M * 128 + block_h_images_bytes(1) + (n_tiles + 1) * 4 + 2 * M * 32 * 4 + 8 * 1024, st));
  void* ring = sc.take(2 * M * 128);
  unsigned char* img = (unsigned char*)sc.take(block_h_images_bytes(1));
  unsigned int* flags = (unsigned int*)sc.take((n_tiles + 1) * 4);
  float* xall = (float*)sc.take(2 * M * 32 * 4);
  if (!xall) return -5;
  RC(split_rows(x, ring, M, st));
  RC(block_h_images(img, filter, gate, dense ? dense : filter, 1, st));      // (the last layer's dense image is never read)
  RC(block_fwd_chain(ring, is_last ? nullptr : xall, zcat, nullptr, ldz, img, prebias, dense_bias, &d, 1, B, T, flags, st,
                     /*ring=*/2, /*last_dense=*/!is_last, /*zcols=*/32));
  if (!is_last) RC((int)cudaMemcpyAsync(x_out, xall + M * 32, (size_t)M * 32 * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int block_bwd_production(const float* x, const float* dx_out, const float* dz_skip, int ldz, float* dx, float* dpre,
                                const float* zcat, const float* filter, const float* gate, const float* dense,
                                const float* prebias, float* gwf, float* gwg, float* gdense, float* gprebias,
                                float* gdense_bias, int B, int T, int d, int is_last, cudaStream_t st) {
  (void)dpre;
  const int64_t M = (int64_t)B * T;
  const int64_t fw = block_bwd_chain_flag_words(1, B, T);
  StreamScratch sc;
  const int64_t pb = bwd_fused_enabled() ? block_bwd_fused_scratch_bytes(1, B, T) : 0;
  RC(sc.init(block_h_images_bytes(1) + block_bwd_h_images_bytes(1) + 2 * M * 32 * 4 + 2 * M * 32 * 2 + 4 * M * 128 + fw * 4 + pb +
             16 * 1024, st));
  unsigned char* img_f = (unsigned char*)sc.take(block_h_images_bytes(1));
  unsigned char* img_b = (unsigned char*)sc.take(block_bwd_h_images_bytes(1));
  float* dzc = (float*)sc.take(M * 32 * 4);      // compact fp32 copies of the pitched inputs
  float* zc = (float*)sc.take(M * 32 * 4);
  void* dz16 = sc.take(M * 32 * 2);
  void* z16 = sc.take(M * 32 * 2);
  void* xs = sc.take(M * 128);
  void* dxs = sc.take(2 * M * 128);               // [dx of this layer | dx' = gradient wrt its output] as split rows
  void* p16 = sc.take(M * 128);
  unsigned int* flags = (unsigned int*)sc.take(fw * 4);
  void* part = pb ? sc.take(pb) : nullptr;
  if (!flags || (pb && !part)) return -5;
  const float* dn = dense ? dense : filter;       // (the last layer's dense images are never read)
  RC(block_h_images(img_f, filter, gate, dn, 1, st));
  RC(block_bwd_h_images(img_b, filter, gate, dn, 1, st));
  RC((int)cudaMemcpy2DAsync(dzc, 32 * 4, dz_skip, (size_t)ldz * 4, 32 * 4, (size_t)M, cudaMemcpyDeviceToDevice, st));
  RC(to_half(dzc, dz16, M * 32, st));
  RC((int)cudaMemcpy2DAsync(zc, 32 * 4, zcat, (size_t)ldz * 4, 32 * 4, (size_t)M, cudaMemcpyDeviceToDevice, st));
  RC(to_half(zc, z16, M * 32, st));
  RC(split_rows(x, xs, M, st));
  if (!is_last) RC(split_rows(dx_out, (char*)dxs + M * 128, M, st));
  if (bwd_fused_enabled()) {
    RC(block_bwd_chain_fused(xs, dxs, p16, dz16, 32, 1.f, img_f, img_b, prebias, &d, 1, B, T, flags, 1.f, gwf, gwg, gdense, gprebias,
                             gdense_bias, part, st, /*last_dense=*/!is_last));
  } else {
    RC(block_bwd_chain(xs, dxs, p16, dz16, 32, 1.f, img_f, img_b, prebias, &d, 1, B, T, flags, st, /*last_dense=*/!is_last));
    RC(block_wgrad_h_all(xs, dxs, p16, z16, 32, 1.f, gwf, gwg, gdense, gprebias, gdense_bias, &d, 1, B, T, st, /*last_dense=*/!is_last));
  }
  RC(unsplit_rows(dxs, dx, M, 1.f, st));
  return 0;
}

// Stand-alone wide block (channels a multiple of 64, R = D): the production kernels of block_wide16.cu as a one-layer
// network.  fp32 in / out at the boundary; inside, the 16-bit storage of the training step (unscaled gradient domain).
static int pitched_to_half(const float* src, int ld, void* dst16, float* tmp, int64_t M, int C_, cudaStream_t st) {
  RC((int)cudaMemcpy2DAsync(tmp, (size_t)C_ * 4, src, (size_t)ld * 4, (size_t)C_ * 4, (size_t)M, cudaMemcpyDeviceToDevice, st));
  return to_half(tmp, dst16, M * C_, st);
}
static int block_fwd_wide(const float* x, float* x_out, float* zcat, int ldz, const float* filter, const float* gate,
                          const float* dense, const float* prebias, const float* dense_bias, int B, int T, int d, int Cw,
                          int is_last, cudaStream_t st) {
  const int64_t M = (int64_t)B * T, n = M * Cw;
  StreamScratch sc;
  RC(sc.init(5 * n * 2 + n * 4 + wide16_images_bytes(1, Cw, Cw) + 16 * 1024, st));
  void* x16 = sc.take(n * 2);
  void* xo16 = sc.take(n * 2);
  void* P16 = sc.take(2 * n * 2);
  void* z16 = sc.take(n * 2);
  float* tmp = (float*)sc.take(n * 4);
  void* img = sc.take(wide16_images_bytes(1, Cw, Cw));
  if (!img) return -5;
  RC(to_half(x, x16, n, st));
  RC(wide16_images(img, filter, gate, dense ? dense : filter, 1, Cw, Cw, st));      // (the last layer's dense images are never read)
  RC(wide16_block_fwd(x16, is_last ? nullptr : xo16, P16, z16, Cw, 0, img, prebias, dense_bias, B, T, d, Cw, Cw, st));
  RC(wide16_to_float(z16, tmp, 1.f, n, st));
  RC((int)cudaMemcpy2DAsync(zcat, (size_t)ldz * 4, tmp, (size_t)Cw * 4, (size_t)Cw * 4, (size_t)M, cudaMemcpyDeviceToDevice, st));
  if (!is_last) RC(wide16_to_float(xo16, x_out, 1.f, n, st));
  return 0;
}
static int block_bwd_wide(const float* x, const float* dx_out, const float* dz_skip, int ldz, float* dx, const float* filter,
                          const float* gate, const float* dense, const float* prebias, float* gwf, float* gwg, float* gdense,
                          float* gprebias, float* gdense_bias, int B, int T, int d, int Cw, int is_last, cudaStream_t st) {
  const int64_t M = (int64_t)B * T, n = M * Cw;
  const int64_t wt = wide16_wgrad_tmp_floats(1, Cw, Cw), csn = (int64_t)B * wide16_colsum_chunks() * 2 * Cw;
  StreamScratch sc;
  RC(sc.init(10 * n * 2 + n * 4 + (wt + csn) * 4 + wide16_images_bytes(1, Cw, Cw) + 32 * 1024, st));
  void* x16 = sc.take(n * 2);
  void* P16 = sc.take(2 * n * 2);
  void* z16 = sc.take(n * 2);
  void* dxn16 = sc.take(n * 2);
  void* dzs16 = sc.take(n * 2);
  void* dz16 = sc.take(n * 2);
  void* dpre16 = sc.take(2 * n * 2);
  void* dxo16 = sc.take(n * 2);
  float* tmp = (float*)sc.take(n * 4);
  float* wtmp = (float*)sc.take(wt * 4);
  float* css = (float*)sc.take(csn * 4);
  void* img = sc.take(wide16_images_bytes(1, Cw, Cw));
  if (!img) return -5;
  RC(to_half(x, x16, n, st));
  RC(wide16_images(img, filter, gate, dense ? dense : filter, 1, Cw, Cw, st));
  // the saved pre-activations (and z) of the forward pass: recomputed here, the entry point is stateless
  RC(wide16_block_fwd(x16, nullptr, P16, z16, Cw, 0, img, prebias, nullptr, B, T, d, Cw, Cw, st));
  RC(pitched_to_half(dz_skip, ldz, dzs16, tmp, M, Cw, st));
  if (!is_last) RC(to_half(dx_out, dxn16, n, st));
  RC((int)cudaMemsetAsync(wtmp, 0, (size_t)wt * 4, st));
  RC(wide16_block_bwd(x16, is_last ? nullptr : dxn16, dzs16, Cw, 0, 1.f, P16, z16, dz16, dpre16, dxo16, img, 1.f, wtmp, gdense,
                      gprebias, nullptr, css, B, T, d, Cw, Cw, st));
  RC(wide16_unpack_wgrad(wtmp, gwf, gwg, 1, Cw, Cw, st));
  if (!is_last && gdense_bias) RC(colsum(dx_out, Cw, (int)M, Cw, gdense_bias, st));      // (in the network: the epilogue of the layer above's dx GEMM)
  return wide16_to_float(dxo16, dx, 1.f, n, st);
}

}  // namespace wn

using namespace wn;

extern "C" {

int wn_abi_version(void) { return WN_ABI_VERSION; }

int wn_debug_set_impl(int32_t gemm_mma, int32_t block_mma) {
  g_gemm_impl = gemm_mma ? 1 : 0;
  set_block_impl(block_mma);
  return 0;
}

int wn_set_grad_ready_event(void* cuda_event) {
  g_tail_ready_event = (cudaEvent_t)cuda_event;
  return 0;
}

int wn_debug_trap_info(unsigned int* host_mapped_words) {
  int rc = block_umma_set_trap_info(host_mapped_words);
  if (rc == 0) rc = block_fwd_h_set_trap_info(host_mapped_words);
  if (rc == 0) rc = block_bwd_h_set_trap_info(host_mapped_words);
  return rc;
}

int wn_debug_timeline(long long* stamps) {
  set_block_timeline(stamps);
  set_fwd_h_timeline(stamps);
  set_bwd_h_timeline(stamps ? stamps + 16 : nullptr);      // slots 16..26: the backward chain's wait cycles
  set_gen_timeline(stamps);
  return 0;
}

int wn_profile_begin(void) {
  if (!g_prof_created) {
    for (int i = 0; i < PROF_MAX; ++i)
      if (cudaEventCreate(&g_prof_ev[i]) != cudaSuccess) return (int)cudaGetLastError();
    g_prof_created = true;
  }
  g_prof_n = 0;
  g_prof_on = true;
  return 0;
}

int wn_profile_mark(int32_t tag, wn_stream_t stream) {
  if (tag < 0 || tag >= PT_COUNT) return -1;
  prof_mark((cudaStream_t)stream, tag);
  return 0;
}

int wn_profile_end(float* ms_per_tag, int32_t* launches_per_tag, int32_t n_tags) {
  g_prof_on = false;
  if (!ms_per_tag || !launches_per_tag || n_tags < PT_COUNT) return -1;
  for (int i = 0; i < n_tags; ++i) { ms_per_tag[i] = 0.f; launches_per_tag[i] = 0; }
  if (g_prof_n == 0) return 0;
  cudaError_t e = cudaEventSynchronize(g_prof_ev[g_prof_n - 1]);
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  for (int i = 1; i < g_prof_n; ++i) {
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, g_prof_ev[i - 1], g_prof_ev[i]);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    ms_per_tag[g_prof_tag[i]] += ms;
    launches_per_tag[g_prof_tag[i]] += 1;
  }
  return 0;
}

int wn_profile_tag_name(int32_t tag, char* out, int32_t n) {
  static const char* names[PT_COUNT] = {
      "misc", "mulaw_encode", "cond_bias_fwd", "frontend_fwd", "block_fwd", "skip_bias_sum", "gemm_skip_fwd",
      "gemm_post1_fwd", "gemm_post2_fwd", "softmax_xent", "gemm_post2_wgrad", "colsum", "gemm_post2_dgrad",
      "gemm_post1_wgrad", "gemm_post1_dgrad", "gemm_skip_wgrad", "gemm_skip_dgrad", "block_bwd_dx",
      "block_wgrad", "frontend_bwd", "cond_bias_bwd", "transpose", "block_bwd_pre"};
  if (tag < 0 || tag >= PT_COUNT || !out || n < 1) return -1;
  snprintf(out, n, "%s", names[tag]);
  return 0;
}

int wn_param_layout(const wn_config* cfg, wn_layout* out) {
  if (!out) return -1;
  return make_layout(cfg, out);
}

int wn_mulaw_encode(const float* audio, int64_t n, const float* thresholds, int32_t q, int32_t* ids,
                    wn_stream_t stream) {
  if (n > 0 && (!audio || !thresholds || !ids)) return -1;
  return mulaw_encode(audio, n, thresholds, q, ids, (cudaStream_t)stream);
}
int wn_mulaw_decode(const int32_t* ids, int64_t n, const float* lut, int32_t q, float* out, wn_stream_t stream) {
  if (n > 0 && (!ids || !lut || !out)) return -1;
  return mulaw_decode(ids, n, lut, q, out, (cudaStream_t)stream);
}

int wn_frontend_fwd(const int32_t* ids, const float* causal_filter, float* x0, int32_t batch, int32_t time,
                    int32_t q, int32_t r, wn_stream_t stream) {
  if (!ids || !causal_filter || !x0 || batch < 1 || time < 1) return -1;
  return frontend_fwd(ids, causal_filter, x0, batch * time, time, q, r, nullptr, (cudaStream_t)stream);
}
int wn_frontend_bwd(const int32_t* ids, const float* dx0, float* grad_causal_filter, int32_t batch, int32_t time,
                    int32_t q, int32_t r, wn_stream_t stream) {
  if (!ids || !dx0 || !grad_causal_filter || batch < 1 || time < 1) return -1;
  return frontend_bwd(ids, dx0, grad_causal_filter, batch * time, time, q, r, (cudaStream_t)stream);
}

int wn_causal_conv(const float* x, const float* w, float* y, int32_t batch, int32_t time, int32_t cin, int32_t cout,
                   int32_t width, int32_t dilation, wn_stream_t stream) {
  if (!x || !w || !y || batch < 1 || time < 1) return -1;
  return causal_conv(x, w, y, batch * time, time, cin, cout, width, dilation, (cudaStream_t)stream);
}

int wn_block_fwd(const float* x, float* x_out, float* zcat, int32_t ldz, const float* filter, const float* gate,
                 const float* dense, const float* prebias, const float* dense_bias, int32_t batch, int32_t time,
                 int32_t dilation, int32_t channels, int32_t is_last, wn_stream_t stream) {
  if (!x || !zcat || !filter || !gate || !prebias || batch < 1 || time < 1 || dilation < 1) return -1;
  if (!is_last && (!x_out || !dense)) return -1;
  if ((ldz & 3) || ldz < channels) return -3;
  if (wide16_supported(channels, channels) && wide16_enabled() && fwd16_enabled())
    return block_fwd_wide(x, x_out, zcat, ldz, filter, gate, dense, prebias, dense_bias, batch, time, dilation, channels, is_last,
                          (cudaStream_t)stream);
  if (channels == 32 && block_umma_enabled() && fwd_h_enabled() && fwd_chain_enabled())
    return block_fwd_production(x, x_out, zcat, ldz, filter, gate, dense, prebias, dense_bias, batch, time, dilation, is_last,
                                (cudaStream_t)stream);
  return block_fwd(x, x_out, zcat, ldz, nullptr, filter, gate, dense, prebias, dense_bias,
                   batch * time, time, dilation, channels, is_last, (cudaStream_t)stream);
}

int wn_block_bwd(const float* x, const float* dx_out, const float* dz_skip, int32_t ldz, float* dx,
                 float* dpre_scratch, const float* zcat, const float* filter, const float* gate, const float* dense,
                 const float* prebias, float* grad_filter, float* grad_gate, float* grad_dense, float* grad_prebias,
                 float* grad_dense_bias, int32_t batch, int32_t time, int32_t dilation, int32_t channels,
                 int32_t is_last, wn_stream_t stream) {
  if (!x || !dz_skip || !dx || !dpre_scratch || !zcat || !filter || !gate || !prebias || !grad_filter ||
      !grad_gate || !grad_prebias || batch < 1 || time < 1 || dilation < 1)
    return -1;
  if (!is_last && (!dx_out || !dense || !grad_dense)) return -1;
  if ((ldz & 3) || ldz < channels) return -3;
  if (wide16_supported(channels, channels) && wide16_enabled() && fwd16_enabled())
    return block_bwd_wide(x, dx_out, dz_skip, ldz, dx, filter, gate, dense, prebias, grad_filter, grad_gate, grad_dense, grad_prebias,
                          grad_dense_bias, batch, time, dilation, channels, is_last, (cudaStream_t)stream);
  if (channels == 32 && block_umma_enabled() && fwd_h_enabled() && bwd_chain_enabled())
    return block_bwd_production(x, dx_out, dz_skip, ldz, dx, dpre_scratch, zcat, filter, gate, dense, prebias, grad_filter,
                                grad_gate, grad_dense, grad_prebias, grad_dense_bias, batch, time, dilation, is_last,
                                (cudaStream_t)stream);
  return block_bwd(x, dx_out, dz_skip, ldz, dx, dpre_scratch, zcat, filter, gate, dense, prebias, grad_filter,
                   grad_gate, grad_dense, grad_prebias, grad_dense_bias, batch * time, time, dilation, channels,
                   is_last, (cudaStream_t)stream);
}

int wn_gemm_tf32(int32_t mode, const float* a, int32_t lda, const float* b, int32_t ldb, float* c, int32_t ldc,
                 int32_t m, int32_t n, int32_t k, const float* bias, const float* relu_mask, int32_t ldmask,
                 int32_t flags, int32_t split_k, wn_stream_t stream) {
  if (!a || !b || !c) return -1;
  GemmParams p = gp(a, lda, b, ldb, c, ldc, m, n, k);
  p.bias = bias;
  p.aux = relu_mask;
  p.ldaux = ldmask;
  p.flags = flags;
  if (mode == 2) p.flags |= GEMM_ATOMIC;
  return gemm_tf32(mode, p, split_k, (cudaStream_t)stream);
}

int wn_gemm_nt_umma(const float* a, int32_t lda, const float* b, int32_t ldb, float* c, int32_t ldc, float* ct,
                    int32_t ldct, int32_t m, int32_t n, int32_t k, const float* bias, const float* relu_mask,
                    int32_t ldmask, int32_t flags, int32_t split_k, wn_stream_t stream) {
  if (!a || !b || (!c && !ct)) return -1;
  GemmParams p = gp(a, lda, b, ldb, c, ldc, m, n, k);
  p.bias = bias;
  p.aux = relu_mask;
  p.ldaux = ldmask;
  p.flags = flags;
  return gemm_nt_umma(p, ct, ldct, split_k, (cudaStream_t)stream);
}

int wn_gemm_f16_nt(const void* a16, int32_t lda, const void* b16, int32_t ldb, float* c, int32_t ldc, void* c16,
                   int32_t ldc16, int32_t m, int32_t n, int32_t k, const float* bias, const float* relu_mask,
                   int32_t ldmask, float c_scale, int32_t flags, wn_stream_t stream) {
  if (!a16 || !b16 || !c) return -1;
  return gemm_f16_nt(a16, lda, b16, ldb, c, ldc, c16, ldc16, m, n, k, bias, relu_mask, ldmask, c_scale, flags,
                     (cudaStream_t)stream);
}

int wn_gemm_f16_nt_colsum(const void* a16, int32_t lda, const void* b16, int32_t ldb, float* c, int32_t ldc, void* c16,
                          int32_t ldc16, int32_t m, int32_t n, int32_t k, float c_scale, float* colsum, float colsum_scale,
                          wn_stream_t stream) {
  if (!a16 || !b16 || !c16 || !colsum) return -1;
  return gemm_f16_nt(a16, lda, b16, ldb, c, ldc, c16, ldc16, m, n, k, nullptr, nullptr, 0, c_scale, 0, (cudaStream_t)stream,
                     nullptr, nullptr, 0, colsum, colsum_scale);
}

int wn_gemm_f16_tn(const void* a16, int32_t lda, const void* b16, int32_t ldb, float* c, int32_t ldc, int32_t m, int32_t n,
                   int32_t k, float c_scale, int32_t split_k, wn_stream_t stream) {
  return gemm_f16_tn(a16, lda, b16, ldb, c, ldc, m, n, k, c_scale, split_k, (cudaStream_t)stream);
}

int wn_gemm_umma(int32_t mode, const float* a, int32_t lda, const float* b, int32_t ldb, float* c, int32_t ldc,
                 int32_t m, int32_t n, int32_t k, const float* bias, const float* relu_mask, int32_t ldmask,
                 int32_t flags, int32_t split_k, wn_stream_t stream) {
  if (!a || !b || !c) return -1;
  GemmParams p = gp(a, lda, b, ldb, c, ldc, m, n, k);
  p.bias = bias;
  p.aux = relu_mask;
  p.ldaux = ldmask;
  p.flags = flags;
  if (mode == 2) p.flags |= GEMM_ATOMIC;
  return gemm_umma(mode, p, nullptr, 0, split_k, (cudaStream_t)stream);
}

int wn_softmax_xent(float* logits, const int32_t* ids, int32_t batch, int32_t time, int32_t q, float* partials,
                    int32_t n_partials, float* loss_out, int32_t write_grad, wn_stream_t stream) {
  if (!logits || !ids || !partials || !loss_out || batch < 1 || time < 1) return -1;
  const int M = batch * time;
  return softmax_xent(logits, ids, M, time, q, 1.0f / (float)M, partials, n_partials, loss_out, write_grad, nullptr, 0.f,
                      (cudaStream_t)stream);
}

int wn_post2_xent(const void* a16, int32_t lda, const void* w16, int32_t ldw, const float* bias, const int32_t* ids,
                  int32_t batch, int32_t time, int32_t k, int32_t q, float* partials, float* loss_out, void* g16,
                  float grad_scale, float* bias_grad, float colsum_scale, wn_stream_t stream) {
  if (!a16 || !w16 || !ids || !partials || !loss_out || !g16 || batch < 1 || time < 1 || lda < k || ldw < k || (lda & 7) || (ldw & 7))
    return -1;
  const int M = batch * time;
  return post2_xent(a16, lda, w16, ldw, bias, ids, M, time, k, q, 1.0f / (float)M, partials, loss_out, g16, grad_scale, bias_grad,
                    colsum_scale, (cudaStream_t)stream);
}

int64_t wn_train_workspace_bytes(const wn_config* cfg, int32_t batch, int32_t time) {
  if (check_cfg(cfg) || batch < 1 || time < 1 || (int64_t)batch * time > (1 << 30)) return -1;
  Workspace w;
  carve(cfg, batch, time, true, nullptr, &w);
  return w.bytes;
}
int64_t wn_forward_workspace_bytes(const wn_config* cfg, int32_t batch, int32_t time) {
  if (check_cfg(cfg) || batch < 1 || time < 1 || (int64_t)batch * time > (1 << 30)) return -1;
  Workspace w;
  carve(cfg, batch, time, false, nullptr, &w);
  return w.bytes;
}

int wn_forward_logits(const wn_config* cfg, const float* params, void* workspace, int64_t workspace_bytes,
                      const int32_t* ids, const int32_t* gc_ids, int32_t batch, int32_t time, float* logits,
                      wn_stream_t stream) {
  wn_layout lo;
  RC(make_layout(cfg, &lo));
  if (!params || !workspace || !ids || !logits || batch < 1 || time < 1) return -1;
  if ((uintptr_t)workspace & 255) return -4;
  Workspace w;
  carve(cfg, batch, time, false, workspace, &w);
  if (w.bytes > workspace_bytes) return -5;
  // (scalar_input: `ids` carries the float32 waveform -- the mu-law decoded samples of predict_proba, model.py:570-576)
  return run_forward(cfg, lo, params, w, ids, gc_ids, batch, time, false, logits, (cudaStream_t)stream,
                     cfg->scalar_input ? reinterpret_cast<const float*>(ids) : nullptr);
}

int wn_predict_last(const wn_config* cfg, const float* params, void* workspace, int64_t workspace_bytes, const int32_t* ids,
                    const int32_t* gc_ids, int32_t batch, int32_t time, float* proba, wn_stream_t stream) {
  wn_layout lo;
  RC(make_layout(cfg, &lo));
  if (!params || !workspace || !ids || !proba || batch < 1 || time < 1) return -1;
  if ((uintptr_t)workspace & 255) return -4;
  Workspace w;
  carve(cfg, batch, time, false, workspace, &w);
  if (w.bytes > workspace_bytes) return -5;
  // the layers run over the whole window, skip sum / postprocess1 / postprocess2 on its last row only; the row of logits
  // goes through the partials scratch (4096 floats)
  if (cfg->quantization_channels > 4096) return -2;
  RC(run_forward(cfg, lo, params, w, ids, gc_ids, batch, time, false, w.partials, (cudaStream_t)stream,
                 cfg->scalar_input ? reinterpret_cast<const float*>(ids) : nullptr, /*last_only=*/true));
  return softmax_f64(w.partials, cfg->quantization_channels, proba, (cudaStream_t)stream);
}

int wn_add_l2(float* loss, const float* params, int64_t n, float coef, wn_stream_t stream) {
  if (!loss || !params || n < 0) return -1;
  return add_l2(loss, params, n, coef, (cudaStream_t)stream);
}

int wn_loss_grad(const wn_config* cfg, const float* params, float* grads, void* workspace, int64_t workspace_bytes,
                 const float* audio, const int32_t* gc_ids, const float* mulaw_thresholds, int32_t batch,
                 int32_t time, float* loss_out, wn_stream_t stream) {
  wn_layout lo;
  RC(make_layout(cfg, &lo));
  if (!params || !grads || !workspace || !audio || !mulaw_thresholds || !loss_out || batch < 1 || time < 1) return -1;
  if ((uintptr_t)workspace & 255) return -4;
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  carve(cfg, batch, time, true, workspace, &w);
  if (w.bytes > workspace_bytes) return -5;
  const int B = batch, T = time, M = B * T, L = cfg->n_layers, R = cfg->residual_channels,
            D = cfg->dilation_channels, S = cfg->skip_channels, Q = cfg->quantization_channels,
            G = cfg->gc_channels;
  const int ldz = L * D;
  const bool rp = cfg->residual_postproc != 0;

  RC((int)cudaMemsetAsync(grads, 0, (size_t)lo.total * sizeof(float), st));
  RC((int)cudaMemsetAsync(w.gprebias, 0, (size_t)L * B * 2 * D * sizeof(float), st));
  prof_mark(st, PT_MISC);
  RC(mulaw_encode(audio, M, mulaw_thresholds, Q, w.ids, st));            // model.py:639
  prof_mark(st, PT_MULAW);
  // postprocess2 + loss + d logits (+ the bias column sums) as one kernel when a logits row fits one accumulator
  // bias gradients (column sums of G1 / G2) collected by the epilogue of the GEMM that produces the matrix (WN_FOLD_COLSUM=0: own launches)
  static const bool fold_env = [] { const char* e = getenv("WN_FOLD_COLSUM"); return !(e && e[0] == '0'); }();
  const bool fold_cs = w.dlog16 && fold_env;
  const bool fused_xent = w.dlog16 && w.Zcat16 && Q == 256 && (S % 64) == 0 && fused_xent_enabled();
  RC(run_forward(cfg, lo, params, w, w.ids, gc_ids, B, T, true, fused_xent ? nullptr : w.logits, st,
                 cfg->scalar_input ? audio : nullptr));      // model.py:645-648
  const int trunc = truncate_stage();
  if (trunc == 1) return 0;
  // fp16 input-gradient chain: gradients travel scaled by gscale = 2^ceil(log2 M), i.e. (softmax - onehot) * [1, 2)
  float gscale = 1.f;
  while (gscale < (float)M) gscale *= 2.f;
  if (fused_xent)
    RC(post2_xent(w.A2h, S, w.W2h, S, P(params, lo.post2_bias), w.ids, M, T, S, Q, 1.0f / (float)M, w.partials, loss_out, w.dlog16,
                  gscale / (float)M, lo.post2_bias >= 0 ? grads + lo.post2_bias : nullptr, 1.f / gscale, st));
  else
    RC(softmax_xent(w.logits, w.ids, M, T, Q, 1.0f / (float)M, w.partials, 4096, loss_out, w.dlog16 ? 0 : 1, w.dlog16, gscale / (float)M, st));      // (fp16 chain: nobody reads the fp32 gradient)
  prof_mark(st, PT_XENT);
  if (trunc == 2) return 0;

  // Weight / bias gradients of the post-processing path only feed the gradient buffers: they run on a second
  // side stream while `st` carries the dependent chain  dlogits -> G1 -> G2 -> dZcat -> residual blocks.
  static cudaStream_t side2 = nullptr;
  static cudaEvent_t ev_g[4] = {nullptr, nullptr, nullptr, nullptr};
  if (!side2) {
    RC((int)cudaStreamCreateWithFlags(&side2, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) RC((int)cudaEventCreateWithFlags(&ev_g[i], cudaEventDisableTiming));
  }
  cudaStream_t s2 = (g_prof_on || no_side_streams() || side_mode() == 2) ? st : side2;   // (per-kernel profiling keeps everything on one stream)
  const float* x2 = rp ? w.T2 : w.A2;   // input of postprocess2
  RC((int)cudaEventRecord(ev_g[0], st));
  RC((int)cudaStreamWaitEvent(s2, ev_g[0], 0));
  {  // postprocess2 gradients:  dW2[S,Q] = X2^T . dlogits
    GemmParams p = gp(x2, S, w.logits, Q, grads + lo.post2, Q, S, Q, M);
    if (w.dlog16)      // fp16 operands: A2 copy and the scaled gradient
      RC(gemm_f16_tn(w.A2h, S, w.dlog16, Q, grads + lo.post2, Q, S, Q, M, 1.f / gscale, split_for(S, Q, M), s2));
    else
      RC(gemm(2, p, split_for(S, Q, M), s2));
    prof_mark(s2, PT_GEMM_POST2_WGRAD);
    if (lo.post2_bias >= 0 && !fused_xent) {
      if (w.dlog16) RC(colsum16(w.dlog16, Q, M, Q, 1.f / gscale, grads + lo.post2_bias, w.cs_scratch, s2));
      else RC(colsum(w.logits, Q, M, Q, grads + lo.post2_bias, s2));
      prof_mark(s2, PT_COLSUM);
    }
  }
  if (w.dlog16) {
    // (fp16 only: every consumer of G1 reads the scaled copy; its column sums = the postprocess1 bias gradient)
    RC(gemm_f16_nt(w.dlog16, Q, w.W2g, Q, nullptr, 0, w.G1h, S, M, S, Q, nullptr, nullptr, 0, 1.f, 0, st, nullptr, w.maskA2, S / 32,
                   fold_cs && lo.post1_bias >= 0 ? grads + lo.post1_bias : nullptr, 1.f / gscale));
    prof_mark(st, PT_GEMM_POST2_DGRAD);
  } else {  // d transformed2 -> d conv1 (relu mask from A2):  G1 = (dlogits . W2^T) * (A2 > 0)
    GemmParams p = gp(w.logits, Q, w.W2R, Q, w.G1, S, M, S, Q);
    p.aux = w.A2; p.ldaux = S;
    p.flags = GEMM_ROUND;
    if (rp) { p.C2 = w.G3; p.ldc2 = S; }
    RC(gemm(1, p, 1, st));
    prof_mark(st, PT_GEMM_POST2_DGRAD);
  }
  RC((int)cudaEventRecord(ev_g[1], st));
  RC((int)cudaStreamWaitEvent(s2, ev_g[1], 0));
  {  // postprocess1 gradients:  dW1[S,S] = A1^T . G1
    GemmParams p = gp(w.A1, S, w.G1, S, grads + lo.post1, S, S, S, M);
    if (w.dlog16)
      RC(gemm_f16_tn(w.A1h, S, w.G1h, S, grads + lo.post1, S, S, S, M, 1.f / gscale, split_for(S, S, M), s2));
    else
      RC(gemm(2, p, split_for(S, S, M), s2));
    prof_mark(s2, PT_GEMM_POST1_WGRAD);
    if (lo.post1_bias >= 0 && !fold_cs) {
      if (w.dlog16) RC(colsum16(w.G1h, S, M, S, 1.f / gscale, grads + lo.post1_bias, w.cs_scratch, s2));
      else RC(colsum(w.G1, S, M, S, grads + lo.post1_bias, s2));
      prof_mark(s2, PT_COLSUM);
    }
  }
  if (w.dlog16) {
    if (fold_cs && lo.skip_bias >= 0) RC((int)cudaMemsetAsync(w.gtmp, 0, S * sizeof(float), st));
    RC(gemm_f16_nt(w.G1h, S, w.W1g, S, nullptr, 0, w.G2h, S, M, S, S, nullptr, nullptr, 0, 1.f, 0, st, nullptr, w.maskA1, S / 32,
                   fold_cs && lo.skip_bias >= 0 ? w.gtmp : nullptr, 1.f / gscale));      // column sums of G2 = every skip bias gradient
    prof_mark(st, PT_GEMM_POST1_DGRAD);
  } else {  // d transformed1 -> d total (relu mask from A1) [+ residual_postproc path]
    GemmParams p = gp(w.G1, S, w.W1R, S, w.G2, S, M, S, S);
    p.aux = w.A1; p.ldaux = S;
    p.flags = rp ? 0 : GEMM_ROUND;
    RC(gemm(1, p, 1, st));
    prof_mark(st, PT_GEMM_POST1_DGRAD);
    if (rp) {
      RC(add_inplace(w.G2, w.G3, (int64_t)M * S, 1, st));
      prof_mark(st, PT_MISC);
    }
  }
  RC((int)cudaEventRecord(ev_g[2], st));
  RC((int)cudaStreamWaitEvent(s2, ev_g[2], 0));
  {  // skip weights / biases:  dWskip[L*D,S] = Zcat^T . G2
    GemmParams p = gp(w.Zcat, ldz, w.G2, S, grads + lo.skip, S, ldz, S, M);
    if (w.dlog16)
      RC(gemm_f16_tn(w.Zcat16, ldz, w.G2h, S, grads + lo.skip, S, ldz, S, M, 1.f / gscale, split_for(ldz, S, M), s2));
    else
      RC(gemm(2, p, split_for(ldz, S, M), s2));
    prof_mark(s2, PT_GEMM_SKIP_WGRAD);
    if (lo.skip_bias >= 0) {
      if (!fold_cs) {
        RC((int)cudaMemsetAsync(w.gtmp, 0, S * sizeof(float), s2));
        if (w.dlog16) RC(colsum16(w.G2h, S, M, S, 1.f / gscale, w.gtmp, w.cs_scratch, s2));
        else RC(colsum(w.G2, S, M, S, w.gtmp, s2));
        prof_mark(s2, PT_COLSUM);
      }
      RC(bcast_rows(w.gtmp, S, grads + lo.skip_bias, L, s2));
      prof_mark(s2, PT_MISC);
    }
  }
  RC((int)cudaEventRecord(ev_g[3], s2));
  if (g_tail_ready_event) RC((int)cudaEventRecord(g_tail_ready_event, s2));      // the all-reduce of the tail bucket may start
  if (w.dlog16) {
    RC(gemm_f16_nt(w.G2h, S, w.Wskipg, S, nullptr, 0, w.dZcat16, ldz, M, ldz, S, nullptr, nullptr, 0, 1.f, 0, st));      // stays fp16 and scaled: block_bwd_pre rescales
    prof_mark(st, PT_GEMM_SKIP_DGRAD);
  } else {  // d z (skip path) for every layer at once:  dZcat = G2 . Wskip^T
    GemmParams p = gp(w.G2, S, w.WskipR, S, w.dZcat, ldz, M, ldz, S);
    RC(gemm(1, p, 1, st));
    prof_mark(st, PT_GEMM_SKIP_DGRAD);
  }
  if (trunc == 3) {
    RC((int)cudaStreamWaitEvent(st, ev_g[3], 0));
    return 0;
  }
  const int64_t xs = (int64_t)M * R;
  const float* dcur = nullptr;     // gradient wrt the output of the layer being processed
  if (w.bwd16) {
    // One persistent, flag-ordered kernel for the pre-activation and input gradients of every layer, then the weight
    // gradients of every layer in one launch; both on fp16 tiles in the domain scaled by gscale * cs (block_bwd_h.cu).
    const float cs = 16.f;
    if (bwd_fused_enabled()) {      // ... or both in ONE launch: the weight gradients from the tiles the chain has in shared memory
      RC(block_bwd_chain_fused(w.XS, w.DXS, w.P16, w.dZcat16, ldz, cs, w.WimgH, w.WimgB, w.prebias, cfg->dilations, L, B, T, w.bflags,
                               1.f / (gscale * cs), grads + lo.filter, grads + lo.gate, grads + lo.dense, w.gprebias,
                               lo.dense_bias >= 0 ? grads + lo.dense_bias : nullptr, w.bpart, st));
      if (trunc == 4) { RC((int)cudaStreamWaitEvent(st, ev_g[3], 0)); return 0; }
    } else {
      RC(block_bwd_chain(w.XS, w.DXS, w.P16, w.dZcat16, ldz, cs, w.WimgH, w.WimgB, w.prebias, cfg->dilations, L, B, T, w.bflags, st));
      if (trunc == 4) { RC((int)cudaStreamWaitEvent(st, ev_g[3], 0)); return 0; }
      RC(block_wgrad_h_all(w.XS, w.DXS, w.P16, w.Zcat16, ldz, 1.f / (gscale * cs), grads + lo.filter, grads + lo.gate,
                           grads + lo.dense, w.gprebias, lo.dense_bias >= 0 ? grads + lo.dense_bias : nullptr, cfg->dilations,
                           L, B, T, st));
    }
    RC(unsplit_rows(w.DXS, w.dX, M, 1.f / (gscale * cs), st));
    prof_mark(st, PT_MISC);
    dcur = w.dX;
  } else if (w.umma_bwd) {
    // Critical chain on `st`: pre(l) -> dx(l) -> pre(l-1) -> ...  The weight-gradient GEMM of a layer only feeds
    // the gradient buffers, so it runs on a side stream, concurrently with the chain.  Every layer has its own dpre
    // and dx buffer: the chain never has to wait for the side stream (a cross-stream wait in front of a
    // programmatically launched kernel proved racy), only the final join does.
    static cudaStream_t side = nullptr;
    static cudaEvent_t ev_pre = nullptr, ev_fork = nullptr, ev_wg[3] = {nullptr, nullptr, nullptr};
    if (!side) {
      RC((int)cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
      RC((int)cudaEventCreateWithFlags(&ev_pre, cudaEventDisableTiming));
      RC((int)cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      for (int i = 0; i < 3; ++i) RC((int)cudaEventCreateWithFlags(&ev_wg[i], cudaEventDisableTiming));
    }
    // (per-kernel profiling serialises everything on one stream so that event deltas are kernel times)
    // Default: the layer weight-gradient kernels stay on the caller's stream.  Running them on the side stream
    // (WN_SIDE_WGRAD=1) hides ~0.15 ms per step but hung graph replays on some B200 boxes (bounded waits
    // trapped after ~60 s; never reproduced with the kernels serialised) -- kept switchable until understood.
    static int side_wgrad = -1;
    if (side_wgrad < 0) side_wgrad = getenv("WN_SIDE_WGRAD") ? 1 : 0;
    cudaStream_t ws = (g_prof_on || no_side_streams() || side_mode() == 3 || !side_wgrad) ? st : side;
    // Default: the weight gradients of all layers in ONE persistent launch behind the pre / dx chain (every layer keeps
    // its own dpre / dx buffers).  WN_WGRAD_PER_LAYER=1: one launch per layer inside the chain, as before.
    static int wgrad_all = -1;
    if (wgrad_all < 0) wgrad_all = (getenv("WN_WGRAD_PER_LAYER") || side_wgrad) ? 0 : 1;
    for (int l = L - 1; l >= 0; --l) {
      const int last = (l == L - 1);
      const unsigned char* img = w.Wimg + (size_t)l * block_img_stride();
      float* dpre = w.dpre + (int64_t)l * M * 2 * D;
      float* dnext = w.dX + (int64_t)l * xs;                       // gradient wrt this layer's input
      // single-stream chain  pre(l) -> dx(l) -> wgrad(l) -> pre(l-1) ...  : programmatic dependent launches (every
      // kernel waits, then triggers iff its successor waits too).  With the weight-gradient kernels on the side
      // stream the chain uses plain launches: events next to programmatic launches proved racy (see below).
      const bool chain_pdl = (ws == st) && bwd_pdl_enabled();
      RC(block_bwd_pre_umma(w.X + l * xs, dcur, w.dZcat, w.dZcat16, 1.f / gscale, ldz, l * D, dpre, img + block_img_off_pre(),
                            w.prebias + (int64_t)l * B * 2 * D, B, T, cfg->dilations[l], last, chain_pdl ? 1 : -1, st));
      RC(block_bwd_dx_umma(dcur, dpre, dnext, img + block_img_off_dx(), B, T, cfg->dilations[l], last,
                           chain_pdl ? ((wgrad_all && l == 0) ? 0 : 1) : -1, st));
      if (wgrad_all) {
        dcur = dnext;
        continue;
      }
      if (ws != st) {
        // The event is recorded AFTER the dx kernel (never between a kernel and a programmatic dependent: such an
        // event was observed to fire when the kernel TRIGGERS, not when it completes).
        RC((int)cudaEventRecord(ev_pre, st));
        RC((int)cudaStreamWaitEvent(ws, ev_pre, 0));
      }
      RC(block_wgrad_umma(w.X + l * xs, dcur, dpre, w.Zcat, ldz, l * D, grads + lo.filter + (int64_t)l * 2 * R * D,
                          grads + lo.gate + (int64_t)l * 2 * R * D, grads + lo.dense + (int64_t)l * D * R,
                          w.gprebias + (int64_t)l * B * 2 * D,
                          lo.dense_bias >= 0 ? grads + lo.dense_bias + (int64_t)l * R : nullptr, B, T, cfg->dilations[l],
                          last, chain_pdl ? (l > 0 ? 1 : 2) : 0, ws));
      if (ws != st) RC((int)cudaEventRecord(ev_wg[l % 3], ws));
      dcur = dnext;
    }
    if (wgrad_all)
      RC(block_wgrad_all(w.X, w.dX, w.dpre, w.Zcat, ldz, grads + lo.filter, grads + lo.gate, grads + lo.dense, w.gprebias,
                         lo.dense_bias >= 0 ? grads + lo.dense_bias : nullptr, cfg->dilations, L, B, T, st));
    // join: the bias / conditioning gradients below read what the weight-gradient kernels accumulated
    if (ws != st)
      for (int i = 0; i < 3 && i < L; ++i) RC((int)cudaStreamWaitEvent(st, ev_wg[i], 0));
  } else if (w.wide16) {
    const float cs = 16.f, inv = 1.f / (gscale * cs);
    RC((int)cudaMemsetAsync(w.wtmp16, 0, (size_t)wide16_wgrad_tmp_floats(L, R, D) * sizeof(float), st));
    const char* dcur16 = nullptr;
    for (int l = L - 1; l >= 0; --l) {
      char* dx_out = (char*)w.dx16w + (int64_t)(l & 1) * xs * 2;
      RC(wide16_block_bwd((char*)w.X16 + (int64_t)l * xs * 2, dcur16, w.dZcat16, ldz, l * D, cs, (char*)w.P16w + (int64_t)l * M * 2 * D * 2,
                          w.Zcat16, w.dz16w, w.dpre16w, dx_out, (char*)w.Wimg16 + (int64_t)l * wide16_images_bytes(1, R, D), inv,
                          w.wtmp16 + (int64_t)l * 2 * R * 2 * D, grads + lo.dense + (int64_t)l * D * R,
                          w.gprebias + (int64_t)l * B * 2 * D, (lo.dense_bias >= 0 && l > 0) ? grads + lo.dense_bias + (int64_t)(l - 1) * R : nullptr,
                          w.cs_scratch2, B, T, cfg->dilations[l], R, D, st));
      dcur16 = dx_out;
    }
    RC(wide16_unpack_wgrad(w.wtmp16, grads + lo.filter, grads + lo.gate, L, R, D, st));
    RC(wide16_to_float(dcur16, w.dX, inv, xs, st));
    prof_mark(st, PT_MISC);
    dcur = w.dX;
  } else if (w.gscratch) {
    float* bufs[2] = {w.dX, w.dX + xs};
    int cur_i = 0;
    for (int l = L - 1; l >= 0; --l) {
      const int last = (l == L - 1);
      float* dnext = bufs[cur_i ^ 1];
      RC(generic_block_bwd(w.X + l * xs, last ? nullptr : dcur, w.dZcat, w.Zcat, ldz, l * D, w.Pall + (int64_t)l * M * 2 * D,
                           dnext, w.dpre, params + lo.filter + (int64_t)l * 2 * R * D, params + lo.gate + (int64_t)l * 2 * R * D,
                           params + lo.dense + (int64_t)l * D * R, grads + lo.filter + (int64_t)l * 2 * R * D,
                           grads + lo.gate + (int64_t)l * 2 * R * D, grads + lo.dense + (int64_t)l * D * R,
                           w.gprebias + (int64_t)l * B * 2 * D,
                           lo.dense_bias >= 0 ? grads + lo.dense_bias + (int64_t)l * R : nullptr, w.gscratch, B, T,
                           cfg->dilations[l], R, D, last, st));
      dcur = dnext;
      cur_i ^= 1;
    }
  } else {
    float* bufs[2] = {w.dX, w.dX + xs};
    int cur_i = 0;
    for (int l = L - 1; l >= 0; --l) {
      const int last = (l == L - 1);
      float* dnext = bufs[cur_i ^ 1];
      RC(block_bwd(w.X + l * xs, last ? nullptr : dcur, w.dZcat + (int64_t)l * D, ldz, dnext, w.dpre,
                   w.Zcat + (int64_t)l * D, params + lo.filter + (int64_t)l * 2 * R * D,
                   params + lo.gate + (int64_t)l * 2 * R * D, params + lo.dense + (int64_t)l * D * R,
                   w.prebias + (int64_t)l * B * 2 * D, grads + lo.filter + (int64_t)l * 2 * R * D,
                   grads + lo.gate + (int64_t)l * 2 * R * D, grads + lo.dense + (int64_t)l * D * R,
                   w.gprebias + (int64_t)l * B * 2 * D,
                   lo.dense_bias >= 0 ? grads + lo.dense_bias + (int64_t)l * R : nullptr, M, T, cfg->dilations[l], R,
                   last, st));
      dcur = dnext;
      cur_i ^= 1;
    }
  }
  RC((int)cudaStreamWaitEvent(st, ev_g[3], 0));    // join the post-processing weight-gradient stream
  if (cfg->scalar_input) RC(scalar_frontend_bwd(audio, dcur, grads + lo.causal, M, T, R, cfg->initial_filter_width, st));
  else RC(frontend_bwd(w.ids, dcur, grads + lo.causal, M, T, Q, R, st));
  prof_mark(st, PT_FRONTEND_BWD);
  if (lo.filter_bias >= 0 || G > 0) {
    RC(cond_bias_bwd(w.gprebias, P(grads, lo.filter_bias), P(grads, lo.gate_bias), P(params, lo.gc_filter),
                     P(params, lo.gc_gate), P(grads, lo.gc_filter), P(grads, lo.gc_gate),
                     P(params, lo.gc_embedding), P(grads, lo.gc_embedding), gc_ids, L, B, D, gc_ids ? G : 0,
                     cfg->gc_cardinality, st));
    prof_mark(st, PT_COND_BIAS_BWD);
  }
  return 0;
}

int wn_optim_adam(float* w, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                  double eps, int64_t step, float l2, float grad_scale, wn_stream_t stream) {
  if (!w || !g || !m || !v || step < 1) return -1;
  const double lr_t = lr * sqrt(1.0 - pow(beta2, (double)step)) / (1.0 - pow(beta1, (double)step));
  return optim_adam(w, g, m, v, n, lr_t, beta1, beta2, eps, l2, grad_scale, (cudaStream_t)stream);
}
int wn_optim_momentum(float* w, const float* g, float* accum, int64_t n, double lr, double momentum, float l2,
                      float grad_scale, wn_stream_t stream) {
  if (!w || !g || !accum) return -1;
  return optim_momentum(w, g, accum, n, lr, momentum, l2, grad_scale, (cudaStream_t)stream);
}
int wn_optim_rmsprop(float* w, const float* g, float* ms, float* mom, int64_t n, double lr, double decay,
                     double momentum, double eps, float l2, float grad_scale, wn_stream_t stream) {
  if (!w || !g || !ms || !mom) return -1;
  return optim_rmsprop(w, g, ms, mom, n, lr, decay, momentum, eps, l2, grad_scale, (cudaStream_t)stream);
}

}  // extern "C"
