// Gated residual block for ANY residual / dilation channel counts (R, D multiples of 4): the reference accepts
// arbitrary widths (wavenet/model.py:236-330, e.g. the scaled 128-channel network of BASELINE config 5), the fused
// tcgen05 block kernels (block_umma.cu / block_fwd_h.cu) exist for R = D = 32.  This path builds a layer from the
// tensor-core GEMM (gemm_umma.cu, mma.sync fallback) plus small fusion kernels:
//   forward   P  = [x[t-d] | x[t]] . [W0 ; W1]          3 split-precision GEMMs (hi.hi + lo.hi + hi.lo), red.add
//             z  = tanh(Pf + b) * sigmoid(Pg + b)         gate kernel (keeps P for the backward pass, writes Zcat)
//             x' = x + bd + z . Wd                        3 split-precision GEMMs into the initialised output
//   backward  dz = dz_skip + dx' . Wd^T ; dpre = [df|dg]  GEMM + kernel
//             dW* = [x[t-d] | x[t]]^T . dpre , z^T . dx'  TN GEMMs straight into the gradient buffers; colsum biases
//             dx = dx' + (dpre . W1^T)[t] + (dpre . W0^T)[t+d]     GEMM + kernel
// Same numeric contract as the fused kernels: split precision forward, single-pass TF32 backward.
#include "common.cuh"
#include "kernels.h"

namespace wn {

namespace {
__device__ __forceinline__ void split_tf32(float v, float& h, float& l) {
  h = round_tf32(v);
  l = round_tf32(v - h);
}

// Xcat[m] = [x[m-d] (zero history per batch element) | x[m]], as tf32 hi / lo parts
__global__ void g_concat_split(const float* __restrict__ x, float* __restrict__ xh, float* __restrict__ xl, int64_t M,
                               int T, int R, int d, int want_lo) {
  const int r4 = R >> 2;
  const int64_t n = M * 2 * r4, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t m = i / (2 * r4);
    const int c = (int)(i % (2 * r4));
    const int t = (int)(m % T);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c >= r4) v = __ldg(reinterpret_cast<const float4*>(x + m * R) + (c - r4));
    else if (t >= d) v = __ldg(reinterpret_cast<const float4*>(x + (m - d) * R) + c);
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    reinterpret_cast<float4*>(xh + m * 2 * R)[c] = h;
    if (want_lo) reinterpret_cast<float4*>(xl + m * 2 * R)[c] = l;
  }
}

// Wcat[k][n] (k < R: past tap, k >= R: current tap; n < D: filter, n >= D: gate) hi / lo;  Wd hi / lo
__global__ void g_weights_split(const float* __restrict__ wf, const float* __restrict__ wg, const float* __restrict__ wd,
                                float* __restrict__ wch, float* __restrict__ wcl, float* __restrict__ wdh,
                                float* __restrict__ wdl, int R, int D) {
  const int n1 = 2 * R * 2 * D, n2 = D * R;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int k = i / (2 * D), n = i % (2 * D);
      const float w = n < D ? wf[(size_t)k * D + n] : wg[(size_t)k * D + (n - D)];   // [tap][R][D] == [2R][D]
      split_tf32(w, wch[i], wcl[i]);
    } else {
      const int j = i - n1;
      split_tf32(wd[j], wdh[j], wdl[j]);
    }
  }
}

// P += prebias (kept for the backward pass); z = tanh(f) sigmoid(g) -> Zcat column block (tf32-rounded) and hi / lo
__global__ void g_gate(float* __restrict__ P, const float* __restrict__ prebias, float* __restrict__ zcat, int ldz,
                       float* __restrict__ zh, float* __restrict__ zl, int64_t M, int T, int D, int want_split) {
  const int64_t n = M * D, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t m = i / D;
    const int j = (int)(i % D);
    const int b = (int)(m / T);
    const float f = P[m * 2 * D + j] + prebias[(size_t)b * 2 * D + j];
    const float g = P[m * 2 * D + D + j] + prebias[(size_t)b * 2 * D + D + j];
    P[m * 2 * D + j] = f;
    P[m * 2 * D + D + j] = g;
    const float z = tanh_f(f) * sigmoid_f(g);
    float h, l;
    split_tf32(z, h, l);
    zcat[m * ldz + j] = h;
    if (want_split) { zh[i] = h; zl[i] = l; }
  }
}

__global__ void g_residual_init(const float* __restrict__ x, const float* __restrict__ bd, float* __restrict__ xout,
                                int64_t M, int R) {
  const int64_t n = M * R, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    xout[i] = x[i] + (bd ? bd[i % R] : 0.f);
}

__global__ void g_copy_cols(const float* __restrict__ src, int lds, float* __restrict__ dst, int64_t M, int D) {
  const int64_t n = M * D, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[(i / D) * lds + (i % D)];
}

// dpre = [df | dg] from the saved pre-activations and dz (tf32-rounded: operand of three GEMMs)
__global__ void g_dpre(const float* __restrict__ P, const float* __restrict__ dz, float* __restrict__ dpre, int64_t M, int D) {
  const int64_t n = M * D, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t m = i / D;
    const int j = (int)(i % D);
    const float tf = tanh_f(P[m * 2 * D + j]), sg = sigmoid_f(P[m * 2 * D + D + j]);
    const float g = dz[i];
    dpre[m * 2 * D + j] = round_tf32(g * sg * (1.f - tf * tf));
    dpre[m * 2 * D + D + j] = round_tf32(g * tf * sg * (1.f - sg));
  }
}

// dx[t] = dx'[t] + T1[t][R:2R] + T1[t+d][0:R]   (T1 = dpre . Wcat^T; the second term stays inside the batch element)
__global__ void g_dx(const float* __restrict__ dxn, const float* __restrict__ t1, float* __restrict__ dx, int64_t M, int T,
                     int R, int d) {
  const int64_t n = M * R, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t m = i / R;
    const int r = (int)(i % R);
    const int t = (int)(m % T);
    float v = t1[m * 2 * R + R + r];
    if (dxn) v += dxn[i];
    if (t + d < T) v += t1[(m + d) * 2 * R + r];
    dx[i] = v;
  }
}

__global__ void g_round(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = round_tf32(in[i]);
}

inline int nblocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = 8LL * sm_count();
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}
GemmParams gpar(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K, int flags) {
  GemmParams p;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.bias = nullptr; p.aux = nullptr; p.ldaux = 0; p.C2 = nullptr; p.ldc2 = 0; p.flags = flags;
  return p;
}
#define GRC(x)             \
  do {                     \
    int rc__ = (x);        \
    if (rc__) return rc__; \
  } while (0)
}  // namespace

int64_t generic_scratch_floats(int64_t M, int R, int D) {
  // Xh, Xl [M,2R] | Zh, Zl [M,D] | Wch, Wcl [2R,2D] | Wdh, Wdl [D,R] | dz [M,D] | T1 [M,2R]
  return 2 * M * 2 * R + 2 * M * D + 2 * (int64_t)2 * R * 2 * D + 2 * (int64_t)D * R + M * D + M * 2 * R + 64;
}

namespace {
struct Scratch {
  float *Xh, *Xl, *Zh, *Zl, *Wch, *Wcl, *Wdh, *Wdl, *dz, *T1;
};
Scratch carve_scratch(float* s, int64_t M, int R, int D) {
  Scratch c;
  c.Xh = s; s += M * 2 * R;
  c.Xl = s; s += M * 2 * R;
  c.Zh = s; s += M * D;
  c.Zl = s; s += M * D;
  c.Wch = s; s += (int64_t)2 * R * 2 * D;
  c.Wcl = s; s += (int64_t)2 * R * 2 * D;
  c.Wdh = s; s += (int64_t)D * R;
  c.Wdl = s; s += (int64_t)D * R;
  c.dz = s; s += M * D;
  c.T1 = s;
  return c;
}
}  // namespace

int generic_block_fwd(const float* x, float* xout, float* zcat, int ldz, int zcol, float* P, const float* wf,
                      const float* wg, const float* dense, const float* prebias, const float* dense_bias,
                      float* scratch, int B, int T, int d, int R, int D, int is_last, cudaStream_t st) {
  const int64_t M = (int64_t)B * T;
  if ((R & 3) || (D & 3) || M > (1 << 30)) return -3;
  Scratch s = carve_scratch(scratch, M, R, D);
  g_concat_split<<<nblocks(M * R / 2), 256, 0, st>>>(x, s.Xh, s.Xl, M, T, R, d, 1);
  g_weights_split<<<nblocks(2 * R * 2 * D + D * R), 256, 0, st>>>(wf, wg, dense, s.Wch, s.Wcl, s.Wdh, s.Wdl, R, D);
  WN_CHECK_LAUNCH();
  GRC((int)cudaMemsetAsync(P, 0, (size_t)M * 2 * D * sizeof(float), st));
  GRC(gemm_dispatch(0, gpar(s.Xh, 2 * R, s.Wch, 2 * D, P, 2 * D, (int)M, 2 * D, 2 * R, GEMM_ATOMIC), 1, st));
  GRC(gemm_dispatch(0, gpar(s.Xl, 2 * R, s.Wch, 2 * D, P, 2 * D, (int)M, 2 * D, 2 * R, GEMM_ATOMIC), 1, st));
  GRC(gemm_dispatch(0, gpar(s.Xh, 2 * R, s.Wcl, 2 * D, P, 2 * D, (int)M, 2 * D, 2 * R, GEMM_ATOMIC), 1, st));
  g_gate<<<nblocks(M * D), 256, 0, st>>>(P, prebias, zcat + zcol, ldz, s.Zh, s.Zl, M, T, D, !is_last);
  WN_CHECK_LAUNCH();
  if (!is_last) {
    g_residual_init<<<nblocks(M * R), 256, 0, st>>>(x, dense_bias, xout, M, R);
    WN_CHECK_LAUNCH();
    GRC(gemm_dispatch(0, gpar(s.Zh, D, s.Wdh, R, xout, R, (int)M, R, D, GEMM_ATOMIC), 1, st));
    GRC(gemm_dispatch(0, gpar(s.Zl, D, s.Wdh, R, xout, R, (int)M, R, D, GEMM_ATOMIC), 1, st));
    GRC(gemm_dispatch(0, gpar(s.Zh, D, s.Wdl, R, xout, R, (int)M, R, D, GEMM_ATOMIC), 1, st));
  }
  prof_mark(st, PT_BLOCK_FWD);
  return 0;
}

int generic_block_bwd(const float* x, const float* dxn, const float* dZcat, const float* zcat, int ldz, int zcol,
                      const float* P, float* dx, float* dpre, const float* wf, const float* wg, const float* dense,
                      float* gwf, float* gwg, float* gdense, float* gprebias, float* gdense_bias, float* scratch,
                      int B, int T, int d, int R, int D, int is_last, cudaStream_t st) {
  const int64_t M = (int64_t)B * T;
  if ((R & 3) || (D & 3)) return -3;
  Scratch s = carve_scratch(scratch, M, R, D);
  g_weights_split<<<nblocks(2 * R * 2 * D + D * R), 256, 0, st>>>(wf, wg, dense, s.Wch, s.Wcl, s.Wdh, s.Wdl, R, D);
  // dz = dz_skip (+ dx' . Wd^T)
  g_copy_cols<<<nblocks(M * D), 256, 0, st>>>(dZcat + zcol, ldz, s.dz, M, D);
  WN_CHECK_LAUNCH();
  const float* dxr = nullptr;
  if (!is_last) {
    g_round<<<nblocks(M * R), 256, 0, st>>>(dxn, s.Zl, M * R);      // tf32-rounded dx' (Zl is free in the backward pass)
    WN_CHECK_LAUNCH();
    dxr = s.Zl;
    GRC(gemm_dispatch(1, gpar(dxr, R, s.Wdh, R, s.dz, D, (int)M, D, R, GEMM_ATOMIC), 1, st));
  }
  g_dpre<<<nblocks(M * D), 256, 0, st>>>(P, s.dz, dpre, M, D);
  g_concat_split<<<nblocks(M * R / 2), 256, 0, st>>>(x, s.Xh, s.Xl, M, T, R, d, 0);
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_BWD_PRE);
  // weight gradients: [x[t-d] | x[t]]^T . [df | dg]  ->  filter / gate [tap][R][D];   z^T . dx' -> dense [D][R]
  const int split = (int)((M / 32 / 8) < 1 ? 1 : ((M / 32 / 8) > 4 * sm_count() ? 4 * sm_count() : (M / 32 / 8)));
  for (int tap = 0; tap < 2; ++tap) {
    GRC(gemm_dispatch(2, gpar(s.Xh + tap * R, 2 * R, dpre, 2 * D, gwf + (size_t)tap * R * D, D, R, D, (int)M, 0), split, st));
    GRC(gemm_dispatch(2, gpar(s.Xh + tap * R, 2 * R, dpre + D, 2 * D, gwg + (size_t)tap * R * D, D, R, D, (int)M, 0), split, st));
  }
  for (int b = 0; b < B; ++b) GRC(colsum(dpre + (size_t)b * T * 2 * D, 2 * D, T, 2 * D, gprebias + (size_t)b * 2 * D, st));
  if (!is_last) {
    GRC(gemm_dispatch(2, gpar(zcat + zcol, ldz, dxr, R, gdense, R, D, R, (int)M, 0), split, st));
    if (gdense_bias) GRC(colsum(dxn, R, (int)M, R, gdense_bias, st));
  }
  prof_mark(st, PT_BLOCK_WGRAD);
  // dx = dx' + (dpre . W1^T)[t] + (dpre . W0^T)[t+d]
  GRC(gemm_dispatch(1, gpar(dpre, 2 * D, s.Wch, 2 * D, s.T1, 2 * R, (int)M, 2 * R, 2 * D, 0), 1, st));
  g_dx<<<nblocks(M * R), 256, 0, st>>>(is_last ? nullptr : dxn, s.T1, dx, M, T, R, d);
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_BWD_DX);
  return 0;
}

}  // namespace wn
