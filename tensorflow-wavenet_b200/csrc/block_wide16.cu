// Gated residual blocks for WIDE channel counts (R, D multiples of 64: BASELINE config 5, R = D = 128) in 16-bit
// storage: wavenet/model.py:236-330 and its autodiff.  Every product runs on the tcgen05 fp16 GEMM (gemm_umma.cu,
// fp32 accumulation in TMEM); dilation is a second TMA box of the SAME activation matrix at row t - d (forward) /
// t + d (backward), zero-filled outside the batch element -- no [x[t-d] | x[t]] concatenation, padding or time_to_batch
// tensor exists.  Residual add, skip-path gradient add and biases are fused into the GEMM epilogues:
//   forward   P16 = [x[t-d] | x[t]] . [W0 ; W1] + b                 two-tap GEMM (K = 2R, N = 2D), fp16 out
//             z   = tanh(Pf) * sigmoid(Pg) -> Zcat16 columns          gate kernel (HBM bound)
//             x'  = x + bd + z . Wd                                   GEMM (K = D, N = R), residual added in the epilogue
//   backward  dz  = cs * dz_skip + dx' . Wd^T                         GEMM (K = R, N = D), skip-path term added in the epilogue
//             dpre = [dz s (1 - t^2) | dz t s (1 - s)]                kernel, from the saved pre-activations P16
//             dW0 = x[t-d]^T . dpre, dW1 = x^T . dpre, dWd = z^T . dx'   fp16 TN GEMMs over time (operands as they lie), biases = column sums
//             dx  = dx' + [dpre[t+d] | dpre[t]] . [W0^T ; W1^T]       two-tap GEMM (K = 4D, N = R), dx' added in the epilogue
// Gradients travel as fp16 in the domain scaled by gscale * cs (see api.cu: gscale = 2^ceil(log2 M) / M).
// The activations between layers are fp16 (north_star config 5 asks for 16-bit storage; fp16 has the bytes of bf16 and
// three more mantissa bits, the values are O(1)); accumulation, biases, parameters and gradients are fp32.
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace wn {

namespace {

// per-layer weight image (halfs): Wc [2D][2R] | Wdt [R][D] | Wdg [D][R] | Wdx [R][4D]
__host__ __device__ inline int64_t img_halfs(int R, int D) { return (int64_t)10 * R * D; }

// filter / gate: [tap][R][D] (tap 0 multiplies x[t-d]);  dense: [D][R]
__global__ void w16_weights_kernel(const float* __restrict__ filter, const float* __restrict__ gate,
                                   const float* __restrict__ dense, __half* __restrict__ img, int L, int R, int D) {
  const int64_t per = img_halfs(R, D);
  const int64_t n = per * L, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int l = (int)(i / per);
    int64_t j = i % per;
    const float* wf = filter + (int64_t)l * 2 * R * D;
    const float* wg = gate + (int64_t)l * 2 * R * D;
    const float* wd = dense + (int64_t)l * D * R;
    float v;
    if (j < (int64_t)4 * R * D) {              // Wc[n][k]: n over [filter D | gate D], k over [past R | current R]
      const int n_ = (int)(j / (2 * R)), k = (int)(j % (2 * R));
      const int tap = k / R, r = k % R;
      v = n_ < D ? wf[((int64_t)tap * R + r) * D + n_] : wg[((int64_t)tap * R + r) * D + (n_ - D)];
    } else if ((j -= (int64_t)4 * R * D) < (int64_t)R * D) {      // Wdt[r][dch] = dense[dch][r]
      const int r = (int)(j / D), dc = (int)(j % D);
      v = wd[(int64_t)dc * R + r];
    } else if ((j -= (int64_t)R * D) < (int64_t)R * D) {          // Wdg = dense as stored
      v = wd[j];
    } else {                                                      // Wdx[r][tap * 2D + n]
      j -= (int64_t)R * D;
      const int r = (int)(j / (4 * D)), c = (int)(j % (4 * D));
      const int tap = c / (2 * D), n_ = c % (2 * D);
      v = n_ < D ? wf[((int64_t)tap * R + r) * D + n_] : wg[((int64_t)tap * R + r) * D + (n_ - D)];
    }
    img[i] = __float2half_rn(v);
  }
}

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 a = __half22float2(h[u]);
    f[2 * u] = a.x; f[2 * u + 1] = a.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    h[u] = __floats2half2_rn(fminf(fmaxf(f[2 * u], -65504.f), 65504.f), fminf(fmaxf(f[2 * u + 1], -65504.f), 65504.f));
  return q;
}

// z = tanh(f) sigmoid(g) from P16 = [f D | g D] (biases already added) -> a D-column block of Zcat16; 8 channels per thread
__global__ void w16_gate_kernel(const __half* __restrict__ P, __half* __restrict__ z, int ldz, int64_t M, int D) {
  const int d8 = D >> 3;
  const int64_t n = M * d8, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t m = i / d8;
    const int c = (int)(i % d8) * 8;
    float f[8], g[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(P + m * 2 * D + c)), f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(P + m * 2 * D + D + c)), g);
#pragma unroll
    for (int u = 0; u < 8; ++u) o[u] = tanh_f(f[u]) * sigmoid_f(g[u]);
    *reinterpret_cast<uint4*>(z + m * ldz + c) = pack8(o);
  }
}

// dpre = [df | dg] from dz (src, times src_scale) and the saved pre-activations; the column sums of dpre (= the gradients
// of the filter / gate biases and of the conditioning projections, per batch element) are collected on the way:
// block (x, b) walks rows x*RS + rsub + k*gridDim.x*RS of batch element b (RS = 256 / (D/8) rows per pass) and writes its
// partial sums to partials[b][x][2D]; w16_colsum_finish_kernel adds them up (atomics on one address serialise).
__global__ void __launch_bounds__(256)
w16_dpre_kernel(const __half* __restrict__ src, int lds, float src_scale, const __half* __restrict__ P,
                __half* __restrict__ dpre, int T, int D, float* __restrict__ partials) {
  extern __shared__ float red[];      // [RS][2D]
  const int d8 = D >> 3, RS = 256 / d8;
  const int cg = threadIdx.x % d8, rsub = threadIdx.x / d8, c = cg * 8;
  const int64_t base = (int64_t)blockIdx.y * T;
  float sf[8], sg[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) sf[u] = sg[u] = 0.f;
  if (rsub < RS)
    for (int t = blockIdx.x * RS + rsub; t < T; t += gridDim.x * RS) {
      const int64_t m = base + t;
      float f[8], g[8], dz[8], df[8], dg[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(P + m * 2 * D + c)), f);
      unpack8(__ldg(reinterpret_cast<const uint4*>(P + m * 2 * D + D + c)), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + m * lds + c)), dz);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float t_ = tanh_f(f[u]), s_ = sigmoid_f(g[u]), q = dz[u] * src_scale;
        df[u] = q * s_ * (1.f - t_ * t_);
        dg[u] = q * t_ * s_ * (1.f - s_);
      }
      const uint4 qf = pack8(df), qg = pack8(dg);
      *reinterpret_cast<uint4*>(dpre + m * 2 * D + c) = qf;
      *reinterpret_cast<uint4*>(dpre + m * 2 * D + D + c) = qg;
      unpack8(qf, df);      // the sums are those of the ROUNDED values (what the weight-gradient GEMMs read)
      unpack8(qg, dg);
#pragma unroll
      for (int u = 0; u < 8; ++u) { sf[u] += df[u]; sg[u] += dg[u]; }
    }
  if (rsub < RS) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      red[rsub * 2 * D + c + u] = sf[u];
      red[rsub * 2 * D + D + c + u] = sg[u];
    }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < 2 * D; n += 256) {
    float t = 0.f;
    for (int r = 0; r < RS; ++r) t += red[r * 2 * D + n];
    partials[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * D + n] = t;
  }
}
// out[b][n] += scale * sum_x partials[b][x][n];  grid (N / 32, B), block (32, 8)
__global__ void w16_colsum_finish_kernel(const float* __restrict__ partials, int chunks, int N, float scale, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const float* pb = partials + (int64_t)blockIdx.y * chunks * N;
  float t = 0.f;
  if (n < N)
    for (int x = threadIdx.y; x < chunks; x += 8) t += pb[(int64_t)x * N + n];
  red[threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][threadIdx.x];
    out[(int64_t)blockIdx.y * N + n] += t * scale;
  }
}

__global__ void w16_to_float_kernel(const __half* __restrict__ in, float* __restrict__ out, float scale, int64_t n8) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(in) + i), f);
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(f[0] * scale, f[1] * scale, f[2] * scale, f[3] * scale);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(f[4] * scale, f[5] * scale, f[6] * scale, f[7] * scale);
  }
}

// tmp [L][tap][R][2D] (filter | gate columns) -> the reference layout filter / gate [L][tap][R][D]
__global__ void w16_unpack_wgrad_kernel(const float* __restrict__ tmp, float* __restrict__ gwf, float* __restrict__ gwg,
                                        int64_t rows, int D) {
  const int64_t n = rows * 2 * D, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / (2 * D);
    const int c = (int)(i % (2 * D));
    if (c < D) gwf[r * D + c] += tmp[i];
    else gwg[r * D + (c - D)] += tmp[i];
  }
}

inline int nblk(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = 16LL * sm_count();
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

#define WRC(x)             \
  do {                     \
    int rc__ = (x);        \
    if (rc__) return rc__; \
  } while (0)

}  // namespace

// (R != D is written for but has no GPU parity case yet: those widths stay on the fp32 GEMM-built blocks, which do)
bool wide16_supported(int R, int D) { return R == D && R >= 64 && !(R & 63) && R <= 256; }
int64_t wide16_images_bytes(int L, int R, int D) { return img_halfs(R, D) * L * 2; }
int64_t wide16_wgrad_tmp_floats(int L, int R, int D) { return (int64_t)L * 2 * R * 2 * D; }
int wide16_colsum_chunks() { return 4 * sm_count(); }      // row chunks per batch element of the dpre kernel (partials: [B][chunks][2D])

int wide16_images(void* img, const float* filter, const float* gate, const float* dense, int L, int R, int D, cudaStream_t st) {
  w16_weights_kernel<<<nblk(img_halfs(R, D) * L), 256, 0, st>>>(filter, gate, dense, (__half*)img, L, R, D);
  WN_CHECK_LAUNCH();
  return 0;
}

int wide16_to_float(const void* in, float* out, float scale, int64_t n, cudaStream_t st) {
  if (n & 7) return -3;
  w16_to_float_kernel<<<nblk(n / 8), 256, 0, st>>>((const __half*)in, out, scale, n / 8);
  WN_CHECK_LAUNCH();
  return 0;
}

// one layer forward: x16 [B*T][R] -> P16 [B*T][2D] (kept for the backward pass), z into Zcat16 columns [zcol, zcol + D),
// x16_out (null for the last layer)
int wide16_block_fwd(const void* x16, void* x16_out, void* P16, void* zcat16, int ldz, int zcol, const void* img_l,
                     const float* prebias, const float* dense_bias, int B, int T, int d, int R, int D, cudaStream_t st) {
  const int64_t M = (int64_t)B * T;
  const __half* img = (const __half*)img_l;
  const __half* Wc = img;
  const __half* Wdt = img + (int64_t)4 * R * D;
  // D = 128: one 128 x 256 accumulator holds [f | g] of 128 steps, the gate runs in the GEMM epilogue (WN_WIDE16_GATE=0: own kernel)
  static const bool gate_env = [] { const char* e = getenv("WN_WIDE16_GATE"); return !(e && e[0] == '0'); }();
  const bool gate_fused = gate_env && D == 128;
  for (int b = 0; b < B; ++b) {      // (per batch element: rows t - d < 0 must read zeros, not the previous element)
    F16Extra ex;
    ex.a_split = R; ex.a_shift = -d;
    if (gate_fused) { ex.gate_z16 = (__half*)zcat16 + zcol + (int64_t)b * T * ldz; ex.ldz = ldz; }      // z from the epilogue
    WRC(gemm_f16_nt((const __half*)x16 + (int64_t)b * T * R, R, Wc, 2 * R, nullptr, 0, (__half*)P16 + (int64_t)b * T * 2 * D, 2 * D, T,
                    2 * D, 2 * R, prebias + (int64_t)b * 2 * D, nullptr, 0, 1.f, 0, st, nullptr, nullptr, 0, nullptr, 0.f, &ex));
  }
  if (!gate_fused) {
    w16_gate_kernel<<<nblk(M * (D >> 3)), 256, 0, st>>>((const __half*)P16, (__half*)zcat16 + zcol, ldz, M, D);
    WN_CHECK_LAUNCH();
  }
  if (x16_out) {
    F16Extra ex;
    ex.aux16 = x16; ex.ldaux16 = R; ex.aux_scale = 1.f;
    WRC(gemm_f16_nt((const __half*)zcat16 + zcol, ldz, Wdt, D, nullptr, 0, x16_out, R, (int)M, R, D, dense_bias, nullptr, 0, 1.f, 0, st,
                    nullptr, nullptr, 0, nullptr, 0.f, &ex));
  }
  prof_mark(st, PT_BLOCK_FWD);
  return 0;
}

// one layer backward.  dxn16: gradient wrt the layer's output (null for the last layer), dz16_skip: the layer's column
// block of dZcat16 (scaled by gscale; cs brings it to the block domain), dx16_out: gradient wrt the layer's input.
// inv_scale = 1 / (gscale * cs).  wtmp: this layer's [2][R][2D] fp32 scratch (zeroed by the caller).  gdense_bias_below: the
// dense-bias gradient of layer l - 1 (column sums of dx; null for layer 0 / no biases).  cs_scratch: B * chunks * 2D floats.
int wide16_block_bwd(const void* x16, const void* dxn16, const void* dzcat16, int ldz, int zcol, float cs, const void* P16,
                     const void* zcat16, void* dz16, void* dpre16, void* dx16_out, const void* img_l, float inv_scale,
                     float* wtmp, float* gdense, float* gprebias, float* gdense_bias_below, float* cs_scratch, int B, int T, int d,
                     int R, int D, cudaStream_t st) {
  const int64_t M = (int64_t)B * T;
  const __half* img = (const __half*)img_l;
  const __half* Wdg = img + (int64_t)5 * R * D;
  const __half* Wdx = img + (int64_t)6 * R * D;
  const __half* dzs = (const __half*)dzcat16 + zcol;
  // D = 128: dpre and its column sums come out of the epilogue of the dz GEMM (WN_WIDE16_DPRE=0: own kernels)
  static const bool dpre_env = [] { const char* e = getenv("WN_WIDE16_DPRE"); return !(e && e[0] == '0'); }();
  const bool dpre_fused = dpre_env && dxn16 && D == 128;
  if (dpre_fused) {
    for (int b = 0; b < B; ++b) {      // (per batch element: the column sums are per element)
      F16Extra ex;
      ex.aux16 = dzs + (int64_t)b * T * ldz; ex.ldaux16 = ldz; ex.aux_scale = cs;
      ex.dpre_P16 = (const __half*)P16 + (int64_t)b * T * 2 * D; ex.ldp = 2 * D;
      WRC(gemm_f16_nt((const __half*)dxn16 + (int64_t)b * T * R, R, Wdg, R, nullptr, 0, (__half*)dpre16 + (int64_t)b * T * 2 * D, 2 * D, T, D, R,
                      nullptr, nullptr, 0, 1.f, 0, st, nullptr, nullptr, 0, gprebias + (int64_t)b * 2 * D, inv_scale, &ex));
    }
  } else if (dxn16) {
    F16Extra ex;
    ex.aux16 = dzs; ex.ldaux16 = ldz; ex.aux_scale = cs;
    WRC(gemm_f16_nt(dxn16, R, Wdg, R, nullptr, 0, dz16, D, (int)M, D, R, nullptr, nullptr, 0, 1.f, 0, st, nullptr, nullptr, 0, nullptr,
                    0.f, &ex));
  }
  if (!dpre_fused) {
    const int RS = 256 / (D >> 3);
    int chunks = (4 * sm_count() + B - 1) / B;
    if (chunks > (T + RS - 1) / RS) chunks = (T + RS - 1) / RS;
    if (chunks > wide16_colsum_chunks()) chunks = wide16_colsum_chunks();
    const size_t sh = (size_t)RS * 2 * D * sizeof(float);
    if (dxn16) w16_dpre_kernel<<<dim3(chunks, B), 256, sh, st>>>((const __half*)dz16, D, 1.f, (const __half*)P16, (__half*)dpre16, T, D, cs_scratch);
    else w16_dpre_kernel<<<dim3(chunks, B), 256, sh, st>>>(dzs, ldz, cs, (const __half*)P16, (__half*)dpre16, T, D, cs_scratch);
    WN_CHECK_LAUNCH();
    w16_colsum_finish_kernel<<<dim3((2 * D + 31) / 32, B), dim3(32, 8), 0, st>>>(cs_scratch, chunks, 2 * D, inv_scale, gprebias);
    WN_CHECK_LAUNCH();
  }
  prof_mark(st, PT_BLOCK_BWD_PRE);
  // weight gradients: contraction over time on the operands as they lie in memory
  const int split = sm_count();
  for (int b = 0; b < B; ++b) {
    const __half* xb = (const __half*)x16 + (int64_t)b * T * R;
    const __half* pb = (const __half*)dpre16 + (int64_t)b * T * 2 * D;
    if ((R & 127) == 0) {      // both taps in one launch: rows [0, R) of wtmp = x[t-d]^T . dpre[t], rows [R, 2R) = x[t]^T . dpre[t]
      WRC(gemm_f16_tn(xb, R, pb, 2 * D, wtmp, 2 * D, 2 * R, 2 * D, T, inv_scale, split / 2, st, R, d));
    } else {
      if (d < T) WRC(gemm_f16_tn(xb, R, pb + (int64_t)d * 2 * D, 2 * D, wtmp, 2 * D, R, 2 * D, T - d, inv_scale, split, st));
      WRC(gemm_f16_tn(xb, R, pb, 2 * D, wtmp + (int64_t)R * 2 * D, 2 * D, R, 2 * D, T, inv_scale, split, st));
    }
  }
  if (dxn16) WRC(gemm_f16_tn((const __half*)zcat16 + zcol, ldz, dxn16, R, gdense, R, D, R, (int)M, inv_scale, split, st));
  prof_mark(st, PT_BLOCK_WGRAD);
  for (int b = 0; b < B; ++b) {
    F16Extra ex;
    ex.a_split = 2 * D; ex.a_shift = d;
    if (dxn16) { ex.aux16 = (const __half*)dxn16 + (int64_t)b * T * R; ex.ldaux16 = R; ex.aux_scale = 1.f; }
    // (the column sums of dx = the dense-bias gradient of the layer below, collected by the epilogue)
    WRC(gemm_f16_nt((const __half*)dpre16 + (int64_t)b * T * 2 * D, 2 * D, Wdx, 4 * D, nullptr, 0, (__half*)dx16_out + (int64_t)b * T * R, R,
                    T, R, 4 * D, nullptr, nullptr, 0, 1.f, 0, st, nullptr, nullptr, 0, gdense_bias_below, gdense_bias_below ? inv_scale : 0.f,
                    &ex));
  }
  prof_mark(st, PT_BLOCK_BWD_DX);
  return 0;
}

int wide16_unpack_wgrad(const float* tmp, float* gwf, float* gwg, int L, int R, int D, cudaStream_t st) {
  const int64_t rows = (int64_t)L * 2 * R;
  w16_unpack_wgrad_kernel<<<nblk(rows * 2 * D), 256, 0, st>>>(tmp, gwf, gwg, rows, D);
  WN_CHECK_LAUNCH();
  return 0;
}

}  // namespace wn
