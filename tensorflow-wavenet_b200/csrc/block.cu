// Gated residual block of the dilated stack: forward, input-gradient and weight-gradient
// kernels (reference: wavenet/model.py:236-330 `_create_dilation_layer`, the two
// `causal_conv` calls of wavenet/ops.py:46-62 and TF autodiff of both).
//
//   pre  = [x[t-d] | x[t]] . Wcat + prebias[b]         Wcat = [[Wf0 Wg0],[Wf1 Wg1]]   (2C x 2C)
//   z    = tanh(pre_f) * sigmoid(pre_g)                 -> Zcat[:, l*C:(l+1)*C]  (skip GEMM input)
//   x'   = x + z . Wd + bd                              (not for the last layer)
//
// Dilation is pure addressing: the "past" operand rows are read at row - d straight from
// the [B*T, C] activation matrix (zero for t < d, per batch element) -- no pad /
// time_to_batch / transpose / batch_to_time / slice tensors exist.
//
// Math runs on the tensor cores as m16n8k8 TF32 MMAs with fp32 accumulation.  A operands
// are loaded straight from global memory as 16-byte vectors: the K index of every MMA is
// permuted (slot t <-> column (C/4)*t+2s, slot t+4 <-> the next column) so that each lane
// owns a contiguous chunk of its row; the weight tiles in shared memory are packed with the
// same permutation as float2 k-pairs with bank-conflict-free pitches.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace wn {

template <int C>
struct BlockCfg {
  static constexpr int NT = 2 * C;                       // width of [f|g] and of [past|cur]
  static constexpr int OFF = (C == 32) ? 1 : 2;          // pitch fix-up (float2 units)
  static constexpr int NTP = NT + OFF;                   // pitch of wcat   (K-perm stride C/8)
  static constexpr int DTP = C + OFF;                    // pitch of wdT    (K-perm stride C/8)
  static constexpr int NDP = C + 4;                      // pitch of wd / wT (K stride 1 per lane t)
  static constexpr int CH = C / 4;                       // floats per lane per row chunk
  static constexpr int KS = C / 8;                       // k-steps per C-wide operand
};

// ---- weight staging -------------------------------------------------------------------
// wcat[kp][n] = (Wcat[2kp][n], Wcat[2kp+1][n]),  Wcat[p][n]: p<C past tap, p>=C current tap;
// n<C filter column, n>=C gate column.  filter/gate are the reference tensors [2][C][C].
template <int C>
__device__ __forceinline__ void stage_wcat(float2* wcat, const float* __restrict__ wf,
                                           const float* __restrict__ wg) {
  using K = BlockCfg<C>;
  for (int i = threadIdx.x; i < C * K::NT; i += blockDim.x) {
    int kp = i / K::NT, n = i % K::NT;
    int p0 = 2 * kp, p1 = p0 + 1;
    const float* w = (n < C) ? wf : wg;
    int col = (n < C) ? n : n - C;
    float a = w[((p0 / C) * C + (p0 % C)) * C + col];
    float b = w[((p1 / C) * C + (p1 % C)) * C + col];
    wcat[kp * K::NTP + n] = make_float2(round_tf32(a), round_tf32(b));
  }
}
// wd[kp][n] = (Wd[2kp][n], Wd[2kp+1][n]),  Wd = dense [C(d)][C(r)]
template <int C>
__device__ __forceinline__ void stage_wd(float2* wd, const float* __restrict__ dense) {
  using K = BlockCfg<C>;
  for (int i = threadIdx.x; i < (C / 2) * C; i += blockDim.x) {
    int kp = i / C, n = i % C;
    wd[kp * K::NDP + n] =
        make_float2(round_tf32(dense[(2 * kp) * C + n]), round_tf32(dense[(2 * kp + 1) * C + n]));
  }
}
// wdT[kp][n] = (Wd[n][2kp], Wd[n][2kp+1])   (k = residual channel, n = dilation channel)
template <int C>
__device__ __forceinline__ void stage_wdT(float2* wdT, const float* __restrict__ dense) {
  using K = BlockCfg<C>;
  for (int i = threadIdx.x; i < (C / 2) * C; i += blockDim.x) {
    int kp = i / C, n = i % C;
    wdT[kp * K::DTP + n] =
        make_float2(round_tf32(dense[n * C + 2 * kp]), round_tf32(dense[n * C + 2 * kp + 1]));
  }
}
// wT_tap[kp][r] = (Wcat[tap*C+r][2kp], Wcat[tap*C+r][2kp+1]),  kp in [0,C), r in [0,C)
template <int C>
__device__ __forceinline__ void stage_wT(float2* wT, const float* __restrict__ wf,
                                         const float* __restrict__ wg, int tap) {
  using K = BlockCfg<C>;
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    int kp = i / C, r = i % C;
    int n0 = 2 * kp, n1 = n0 + 1;
    float a = (n0 < C) ? wf[(tap * C + r) * C + n0] : wg[(tap * C + r) * C + (n0 - C)];
    float b = (n1 < C) ? wf[(tap * C + r) * C + n1] : wg[(tap * C + r) * C + (n1 - C)];
    wT[kp * K::NDP + r] = make_float2(round_tf32(a), round_tf32(b));
  }
}

// ---- fragment helpers -------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void load_chunk(uint32_t (&v)[C / 4], const float* __restrict__ row, int t,
                                           bool valid) {
  if (valid) {
    const float4* p = reinterpret_cast<const float4*>(row + (C / 4) * t);
#pragma unroll
    for (int i = 0; i < C / 16; ++i) {
      float4 q = __ldg(p + i);
      v[4 * i + 0] = f2tf32(q.x);
      v[4 * i + 1] = f2tf32(q.y);
      v[4 * i + 2] = f2tf32(q.z);
      v[4 * i + 3] = f2tf32(q.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) v[i] = 0u;
  }
}

// acc[j][.] (j over 2C/8 n-tiles) += [past | cur] . Wcat
template <int C>
__device__ __forceinline__ void mma_pre(float (&acc)[2 * C / 8][4], const uint32_t (&p0)[C / 4],
                                        const uint32_t (&p1)[C / 4], const uint32_t (&c0)[C / 4],
                                        const uint32_t (&c1)[C / 4], const float2* __restrict__ wcat,
                                        int g, int t) {
  using K = BlockCfg<C>;
#pragma unroll
  for (int part = 0; part < 2; ++part) {
#pragma unroll
    for (int s = 0; s < K::KS; ++s) {
      uint32_t a0 = part ? c0[2 * s] : p0[2 * s];
      uint32_t a1 = part ? c1[2 * s] : p1[2 * s];
      uint32_t a2 = part ? c0[2 * s + 1] : p0[2 * s + 1];
      uint32_t a3 = part ? c1[2 * s + 1] : p1[2 * s + 1];
      int kp = part * (C / 2) + K::KS * t + s;
      const float2* wrow = wcat + kp * K::NTP + g;
#pragma unroll
      for (int j = 0; j < K::NT / 8; ++j) {
        float2 b = wrow[8 * j];
        mma_tf32(acc[j], a0, a1, a2, a3, __float_as_uint(b.x), __float_as_uint(b.y));
      }
    }
  }
}

// =========================================================================================
// forward
// =========================================================================================
template <int C>
__global__ void __launch_bounds__(256, 2)
block_fwd_kernel(const float* __restrict__ x, float* __restrict__ xout, float* __restrict__ zc, int ldz,
                 const float* __restrict__ wf, const float* __restrict__ wg,
                 const float* __restrict__ dense, const float* __restrict__ prebias,
                 const float* __restrict__ dense_bias, int M, int T, int d, int is_last) {
  using K = BlockCfg<C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* wcat = reinterpret_cast<float2*>(smem_raw);
  float2* wd = wcat + C * K::NTP;
  stage_wcat<C>(wcat, wf, wg);
  if (!is_last) stage_wd<C>(wd, dense);
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warps_per_cta = blockDim.x >> 5;
  const int n_tiles = (M + 15) >> 4;
  for (int tile = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); tile < n_tiles;
       tile += gridDim.x * warps_per_cta) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < M, v1 = r1 < M;
    const int b0 = v0 ? r0 / T : 0, b1 = v1 ? r1 / T : 0;
    const int t0 = r0 - b0 * T, t1 = r1 - b1 * T;
    uint32_t p0[K::CH], p1[K::CH], c0[K::CH], c1[K::CH];
    load_chunk<C>(c0, x + (size_t)r0 * C, t, v0);
    load_chunk<C>(c1, x + (size_t)r1 * C, t, v1);
    load_chunk<C>(p0, x + (size_t)(r0 - d) * C, t, v0 && t0 >= d);
    load_chunk<C>(p1, x + (size_t)(r1 - d) * C, t, v1 && t1 >= d);

    float acc[K::NT / 8][4];
#pragma unroll
    for (int j = 0; j < K::NT / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    mma_pre<C>(acc, p0, p1, c0, c1, wcat, g, t);

    // gated activation in the accumulator layout: (g, 8j+2t), (g, 8j+2t+1), (g+8, .), (g+8, .+1)
    const float* pb0 = prebias + (size_t)b0 * K::NT;
    const float* pb1 = prebias + (size_t)b1 * K::NT;
    uint32_t za[K::KS][4];
#pragma unroll
    for (int j = 0; j < K::KS; ++j) {
      const int col = 8 * j + 2 * t;
      float2 bf0 = __ldg(reinterpret_cast<const float2*>(pb0 + col));
      float2 bg0 = __ldg(reinterpret_cast<const float2*>(pb0 + C + col));
      float2 bf1 = __ldg(reinterpret_cast<const float2*>(pb1 + col));
      float2 bg1 = __ldg(reinterpret_cast<const float2*>(pb1 + C + col));
      float z00 = round_tf32(tanh_f(acc[j][0] + bf0.x) * sigmoid_f(acc[j + K::KS][0] + bg0.x));
      float z01 = round_tf32(tanh_f(acc[j][1] + bf0.y) * sigmoid_f(acc[j + K::KS][1] + bg0.y));
      float z10 = round_tf32(tanh_f(acc[j][2] + bf1.x) * sigmoid_f(acc[j + K::KS][2] + bg1.x));
      float z11 = round_tf32(tanh_f(acc[j][3] + bf1.y) * sigmoid_f(acc[j + K::KS][3] + bg1.y));
      if (v0) *reinterpret_cast<float2*>(zc + (size_t)r0 * ldz + col) = make_float2(z00, z01);
      if (v1) *reinterpret_cast<float2*>(zc + (size_t)r1 * ldz + col) = make_float2(z10, z11);
      za[j][0] = __float_as_uint(z00);
      za[j][1] = __float_as_uint(z10);
      za[j][2] = __float_as_uint(z01);
      za[j][3] = __float_as_uint(z11);
    }
    if (is_last) continue;

    float o[K::KS][4];
#pragma unroll
    for (int i = 0; i < K::KS; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
    for (int j = 0; j < K::KS; ++j) {
      const float2* wrow = wd + (4 * j + t) * K::NDP + g;
#pragma unroll
      for (int i = 0; i < K::KS; ++i) {
        float2 b = wrow[8 * i];
        mma_tf32(o[i], za[j][0], za[j][1], za[j][2], za[j][3], __float_as_uint(b.x), __float_as_uint(b.y));
      }
    }
#pragma unroll
    for (int i = 0; i < K::KS; ++i) {
      const int col = 8 * i + 2 * t;
      float2 bd = make_float2(0.f, 0.f);
      if (dense_bias) bd = __ldg(reinterpret_cast<const float2*>(dense_bias + col));
      if (v0) {
        float2 xi = __ldg(reinterpret_cast<const float2*>(x + (size_t)r0 * C + col));
        *reinterpret_cast<float2*>(xout + (size_t)r0 * C + col) =
            make_float2(xi.x + o[i][0] + bd.x, xi.y + o[i][1] + bd.y);
      }
      if (v1) {
        float2 xi = __ldg(reinterpret_cast<const float2*>(x + (size_t)r1 * C + col));
        *reinterpret_cast<float2*>(xout + (size_t)r1 * C + col) =
            make_float2(xi.x + o[i][2] + bd.x, xi.y + o[i][3] + bd.y);
      }
    }
  }
}


// ---- split-precision ("3xTF32") forward ---------------------------------------------------
// A single TF32 pass rounds both operands to 11 bits; through 50 residual layers that error
// accumulates in the residual stream and alone pushes the logits ~2e-3 away from the fp32
// reference (measured).  The forward block therefore splits every operand
// into hi + lo TF32 parts and issues a.lo*b.hi + a.hi*b.lo + a.hi*b.hi (fp32-grade products).
template <int C>
__device__ __forceinline__ void load_chunk3(uint32_t (&hi)[C / 4], uint32_t (&lo)[C / 4],
                                            const float* __restrict__ row, int t, bool valid) {
  if (valid) {
    const float4* p = reinterpret_cast<const float4*>(row + (C / 4) * t);
#pragma unroll
    for (int i = 0; i < C / 16; ++i) {
      float4 q = __ldg(p + i);
      const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t h = f2tf32(v[e]);
        hi[4 * i + e] = h;
        lo[4 * i + e] = f2tf32(v[e] - __uint_as_float(h));
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < C / 4; ++i) hi[i] = lo[i] = 0u;
  }
}

template <int C>
__device__ __forceinline__ void split_tile(float2* hi, float2* lo, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float2 w = lo[i];   // staged as raw fp32 in `lo`
    const float hx = round_tf32(w.x), hy = round_tf32(w.y);
    hi[i] = make_float2(hx, hy);
    lo[i] = make_float2(round_tf32(w.x - hx), round_tf32(w.y - hy));
  }
}

template <int C>
__global__ void __launch_bounds__(256, 2)
block_fwd3_kernel(const float* __restrict__ x, float* __restrict__ xout, float* __restrict__ zc, int ldz,
                  const float* __restrict__ wf, const float* __restrict__ wg,
                  const float* __restrict__ dense, const float* __restrict__ prebias,
                  const float* __restrict__ dense_bias, int M, int T, int d, int is_last) {
  using K = BlockCfg<C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* wcat_hi = reinterpret_cast<float2*>(smem_raw);
  float2* wcat_lo = wcat_hi + C * K::NTP;
  float2* wd_hi = wcat_lo + C * K::NTP;
  float2* wd_lo = wd_hi + (C / 2) * K::NDP;
  // stage raw fp32 pairs into the *_lo tiles, then split in place
  for (int i = threadIdx.x; i < C * K::NT; i += blockDim.x) {
    const int kp = i / K::NT, n = i % K::NT;
    const float* w = (n < C) ? wf : wg;
    const int col = (n < C) ? n : n - C;
    wcat_lo[kp * K::NTP + n] = make_float2(w[(2 * kp) * C + col], w[(2 * kp + 1) * C + col]);
  }
  if (!is_last)
    for (int i = threadIdx.x; i < (C / 2) * C; i += blockDim.x) {
      const int kp = i / C, n = i % C;
      wd_lo[kp * K::NDP + n] = make_float2(dense[(2 * kp) * C + n], dense[(2 * kp + 1) * C + n]);
    }
  __syncthreads();
  split_tile<C>(wcat_hi, wcat_lo, C * K::NTP);
  if (!is_last) split_tile<C>(wd_hi, wd_lo, (C / 2) * K::NDP);
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warps_per_cta = blockDim.x >> 5;
  const int n_tiles = (M + 15) >> 4;
  for (int tile = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); tile < n_tiles;
       tile += gridDim.x * warps_per_cta) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < M, v1 = r1 < M;
    const int b0 = v0 ? r0 / T : 0, b1 = v1 ? r1 / T : 0;
    const int t0 = r0 - b0 * T, t1 = r1 - b1 * T;
    float acc[K::NT / 8][4];
#pragma unroll
    for (int j = 0; j < K::NT / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      uint32_t h0[K::CH], l0[K::CH], h1[K::CH], l1[K::CH];
      if (part == 0) {
        load_chunk3<C>(h0, l0, x + (size_t)(r0 - d) * C, t, v0 && t0 >= d);
        load_chunk3<C>(h1, l1, x + (size_t)(r1 - d) * C, t, v1 && t1 >= d);
      } else {
        load_chunk3<C>(h0, l0, x + (size_t)r0 * C, t, v0);
        load_chunk3<C>(h1, l1, x + (size_t)r1 * C, t, v1);
      }
#pragma unroll
      for (int s = 0; s < K::KS; ++s) {
        const int kp = part * (C / 2) + K::KS * t + s;
        const float2* whi = wcat_hi + kp * K::NTP + g;
        const float2* wlo = wcat_lo + kp * K::NTP + g;
#pragma unroll
        for (int j = 0; j < K::NT / 8; ++j) {
          const float2 bh = whi[8 * j], bl = wlo[8 * j];
          mma_tf32(acc[j], l0[2 * s], l1[2 * s], l0[2 * s + 1], l1[2 * s + 1], __float_as_uint(bh.x),
                   __float_as_uint(bh.y));
          mma_tf32(acc[j], h0[2 * s], h1[2 * s], h0[2 * s + 1], h1[2 * s + 1], __float_as_uint(bl.x),
                   __float_as_uint(bl.y));
          mma_tf32(acc[j], h0[2 * s], h1[2 * s], h0[2 * s + 1], h1[2 * s + 1], __float_as_uint(bh.x),
                   __float_as_uint(bh.y));
        }
      }
    }
    const float* pb0 = prebias + (size_t)b0 * K::NT;
    const float* pb1 = prebias + (size_t)b1 * K::NT;
    uint32_t zh[K::KS][4], zl[K::KS][4];
#pragma unroll
    for (int j = 0; j < K::KS; ++j) {
      const int col = 8 * j + 2 * t;
      float2 bf0 = __ldg(reinterpret_cast<const float2*>(pb0 + col));
      float2 bg0 = __ldg(reinterpret_cast<const float2*>(pb0 + C + col));
      float2 bf1 = __ldg(reinterpret_cast<const float2*>(pb1 + col));
      float2 bg1 = __ldg(reinterpret_cast<const float2*>(pb1 + C + col));
      float z[4];
      z[0] = tanhf(acc[j][0] + bf0.x) * (1.0f / (1.0f + expf(-(acc[j + K::KS][0] + bg0.x))));
      z[1] = tanhf(acc[j][1] + bf0.y) * (1.0f / (1.0f + expf(-(acc[j + K::KS][1] + bg0.y))));
      z[2] = tanhf(acc[j][2] + bf1.x) * (1.0f / (1.0f + expf(-(acc[j + K::KS][2] + bg1.x))));
      z[3] = tanhf(acc[j][3] + bf1.y) * (1.0f / (1.0f + expf(-(acc[j + K::KS][3] + bg1.y))));
      uint32_t h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = f2tf32(z[e]);
        l[e] = f2tf32(z[e] - __uint_as_float(h[e]));
      }
      // Zcat feeds the single-pass TF32 skip GEMM: store the tf32-rounded value
      if (v0) *reinterpret_cast<float2*>(zc + (size_t)r0 * ldz + col) =
          make_float2(__uint_as_float(h[0]), __uint_as_float(h[1]));
      if (v1) *reinterpret_cast<float2*>(zc + (size_t)r1 * ldz + col) =
          make_float2(__uint_as_float(h[2]), __uint_as_float(h[3]));
      zh[j][0] = h[0]; zh[j][1] = h[2]; zh[j][2] = h[1]; zh[j][3] = h[3];
      zl[j][0] = l[0]; zl[j][1] = l[2]; zl[j][2] = l[1]; zl[j][3] = l[3];
    }
    if (is_last) continue;
    float o[K::KS][4];
#pragma unroll
    for (int i = 0; i < K::KS; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
    for (int j = 0; j < K::KS; ++j) {
      const float2* whi = wd_hi + (4 * j + t) * K::NDP + g;
      const float2* wlo = wd_lo + (4 * j + t) * K::NDP + g;
#pragma unroll
      for (int i = 0; i < K::KS; ++i) {
        const float2 bh = whi[8 * i], bl = wlo[8 * i];
        mma_tf32(o[i], zl[j][0], zl[j][1], zl[j][2], zl[j][3], __float_as_uint(bh.x), __float_as_uint(bh.y));
        mma_tf32(o[i], zh[j][0], zh[j][1], zh[j][2], zh[j][3], __float_as_uint(bl.x), __float_as_uint(bl.y));
        mma_tf32(o[i], zh[j][0], zh[j][1], zh[j][2], zh[j][3], __float_as_uint(bh.x), __float_as_uint(bh.y));
      }
    }
#pragma unroll
    for (int i = 0; i < K::KS; ++i) {
      const int col = 8 * i + 2 * t;
      float2 bd = make_float2(0.f, 0.f);
      if (dense_bias) bd = __ldg(reinterpret_cast<const float2*>(dense_bias + col));
      if (v0) {
        float2 xi = __ldg(reinterpret_cast<const float2*>(x + (size_t)r0 * C + col));
        *reinterpret_cast<float2*>(xout + (size_t)r0 * C + col) =
            make_float2(xi.x + o[i][0] + bd.x, xi.y + o[i][1] + bd.y);
      }
      if (v1) {
        float2 xi = __ldg(reinterpret_cast<const float2*>(x + (size_t)r1 * C + col));
        *reinterpret_cast<float2*>(xout + (size_t)r1 * C + col) =
            make_float2(xi.x + o[i][2] + bd.x, xi.y + o[i][3] + bd.y);
      }
    }
  }
}

// =========================================================================================
// backward, input gradient:  dx[m] = dxn[m] + dpre[m].Wcur^T + dpre[m+d].Wpast^T
// dpre = [df | dg] is recomputed for both row sets from x (no activations besides x and
// Zcat's gradient are read); the unshifted dpre is also written (tf32-rounded) to `dpre_out`
// for the weight-gradient kernel.
// =========================================================================================
template <int C>
struct DpreOut {
  uint32_t a[2 * C / 8][4];  // A-fragments of [df|dg] for k-step jj: (g,c) (g+8,c) (g,c+1) (g+8,c+1)
};

template <int C>
__device__ __forceinline__ void compute_dpre(DpreOut<C>& out, const float* __restrict__ x,
                                             const float* __restrict__ dxn, const float* __restrict__ dzs,
                                             int ldz, const float* __restrict__ prebias,
                                             const float2* __restrict__ wcat, const float2* __restrict__ wdT,
                                             int r0, int r1, bool v0, bool v1, int T, int d, int is_last,
                                             int g, int t) {
  using K = BlockCfg<C>;
  const int b0 = v0 ? r0 / T : 0, b1 = v1 ? r1 / T : 0;
  const int t0 = r0 - b0 * T, t1 = r1 - b1 * T;
  float acc[K::NT / 8][4];
#pragma unroll
  for (int j = 0; j < K::NT / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  {
    uint32_t p0[K::CH], p1[K::CH], c0[K::CH], c1[K::CH];
    load_chunk<C>(c0, x + (size_t)r0 * C, t, v0);
    load_chunk<C>(c1, x + (size_t)r1 * C, t, v1);
    load_chunk<C>(p0, x + (size_t)(r0 - d) * C, t, v0 && t0 >= d);
    load_chunk<C>(p1, x + (size_t)(r1 - d) * C, t, v1 && t1 >= d);
    mma_pre<C>(acc, p0, p1, c0, c1, wcat, g, t);
  }
  // dz = dzs + dxn . Wd^T
  float dz[K::KS][4];
#pragma unroll
  for (int j = 0; j < K::KS; ++j) {
    const int col = 8 * j + 2 * t;
    float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
    if (v0) a = __ldg(reinterpret_cast<const float2*>(dzs + (size_t)r0 * ldz + col));
    if (v1) b = __ldg(reinterpret_cast<const float2*>(dzs + (size_t)r1 * ldz + col));
    dz[j][0] = a.x; dz[j][1] = a.y; dz[j][2] = b.x; dz[j][3] = b.y;
  }
  if (!is_last) {
    uint32_t n0[K::CH], n1[K::CH];
    load_chunk<C>(n0, dxn + (size_t)r0 * C, t, v0);
    load_chunk<C>(n1, dxn + (size_t)r1 * C, t, v1);
#pragma unroll
    for (int s = 0; s < K::KS; ++s) {
      const float2* wrow = wdT + (K::KS * t + s) * K::DTP + g;
#pragma unroll
      for (int j = 0; j < K::KS; ++j) {
        float2 b = wrow[8 * j];
        mma_tf32(dz[j], n0[2 * s], n1[2 * s], n0[2 * s + 1], n1[2 * s + 1], __float_as_uint(b.x),
                 __float_as_uint(b.y));
      }
    }
  }
  const float* pb0 = prebias + (size_t)b0 * K::NT;
  const float* pb1 = prebias + (size_t)b1 * K::NT;
#pragma unroll
  for (int j = 0; j < K::KS; ++j) {
    const int col = 8 * j + 2 * t;
    float2 bf0 = __ldg(reinterpret_cast<const float2*>(pb0 + col));
    float2 bg0 = __ldg(reinterpret_cast<const float2*>(pb0 + C + col));
    float2 bf1 = __ldg(reinterpret_cast<const float2*>(pb1 + col));
    float2 bg1 = __ldg(reinterpret_cast<const float2*>(pb1 + C + col));
    const float bfv[4] = {bf0.x, bf0.y, bf1.x, bf1.y};
    const float bgv[4] = {bg0.x, bg0.y, bg1.x, bg1.y};
    float df[4], dg[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool v = (e < 2) ? v0 : v1;
      float tf = tanh_f(acc[j][e] + bfv[e]);
      float sg = sigmoid_f(acc[j + K::KS][e] + bgv[e]);
      float dzv = v ? dz[j][e] : 0.f;
      df[e] = round_tf32(dzv * sg * (1.f - tf * tf));
      dg[e] = round_tf32(dzv * tf * sg * (1.f - sg));
    }
    out.a[j][0] = __float_as_uint(df[0]);
    out.a[j][1] = __float_as_uint(df[2]);
    out.a[j][2] = __float_as_uint(df[1]);
    out.a[j][3] = __float_as_uint(df[3]);
    out.a[j + K::KS][0] = __float_as_uint(dg[0]);
    out.a[j + K::KS][1] = __float_as_uint(dg[2]);
    out.a[j + K::KS][2] = __float_as_uint(dg[1]);
    out.a[j + K::KS][3] = __float_as_uint(dg[3]);
  }
}

template <int C>
__global__ void __launch_bounds__(256, 1)
block_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dxn,
                    const float* __restrict__ dzs, int ldz, float* __restrict__ dx,
                    float* __restrict__ dpre_out, const float* __restrict__ wf,
                    const float* __restrict__ wg, const float* __restrict__ dense,
                    const float* __restrict__ prebias, int M, int T, int d, int is_last) {
  using K = BlockCfg<C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* wcat = reinterpret_cast<float2*>(smem_raw);
  float2* wdT = wcat + C * K::NTP;
  float2* wTcur = wdT + (C / 2) * K::DTP;
  float2* wTpast = wTcur + C * K::NDP;
  stage_wcat<C>(wcat, wf, wg);
  if (!is_last) stage_wdT<C>(wdT, dense);
  stage_wT<C>(wTcur, wf, wg, 1);
  stage_wT<C>(wTpast, wf, wg, 0);
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warps_per_cta = blockDim.x >> 5;
  const int n_tiles = (M + 15) >> 4;
  for (int tile = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); tile < n_tiles;
       tile += gridDim.x * warps_per_cta) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < M, v1 = r1 < M;
    float o[K::KS][4];
#pragma unroll
    for (int i = 0; i < K::KS; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

#pragma unroll 1
    for (int sub = 0; sub < 2; ++sub) {
      // sub 0: rows m (current-tap path);  sub 1: rows m+d (past-tap path, same batch element only)
      int q0 = r0, q1 = r1;
      bool w0 = v0, w1 = v1;
      if (sub == 1) {
        const int t0 = v0 ? r0 % T : 0, t1 = v1 ? r1 % T : 0;
        w0 = v0 && (t0 + d < T);
        w1 = v1 && (t1 + d < T);
        q0 = r0 + d;
        q1 = r1 + d;
        if (!__any_sync(0xffffffffu, w0 || w1)) break;
      }
      DpreOut<C> dp;
      compute_dpre<C>(dp, x, dxn, dzs, ldz, prebias, wcat, wdT, q0, q1, w0, w1, T, d, is_last, g, t);
      if (sub == 0) {
#pragma unroll
        for (int jj = 0; jj < K::NT / 8; ++jj) {
          const int col = 8 * jj + 2 * t;
          if (v0)
            *reinterpret_cast<float2*>(dpre_out + (size_t)r0 * K::NT + col) =
                make_float2(__uint_as_float(dp.a[jj][0]), __uint_as_float(dp.a[jj][2]));
          if (v1)
            *reinterpret_cast<float2*>(dpre_out + (size_t)r1 * K::NT + col) =
                make_float2(__uint_as_float(dp.a[jj][1]), __uint_as_float(dp.a[jj][3]));
        }
      }
      const float2* wT = sub ? wTpast : wTcur;
#pragma unroll
      for (int jj = 0; jj < K::NT / 8; ++jj) {
        const float2* wrow = wT + (4 * jj + t) * K::NDP + g;
#pragma unroll
        for (int i = 0; i < K::KS; ++i) {
          float2 b = wrow[8 * i];
          mma_tf32(o[i], dp.a[jj][0], dp.a[jj][1], dp.a[jj][2], dp.a[jj][3], __float_as_uint(b.x),
                   __float_as_uint(b.y));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < K::KS; ++i) {
      const int col = 8 * i + 2 * t;
      if (v0) {
        float2 a = make_float2(0.f, 0.f);
        if (!is_last) a = __ldg(reinterpret_cast<const float2*>(dxn + (size_t)r0 * C + col));
        *reinterpret_cast<float2*>(dx + (size_t)r0 * C + col) = make_float2(a.x + o[i][0], a.y + o[i][1]);
      }
      if (v1) {
        float2 a = make_float2(0.f, 0.f);
        if (!is_last) a = __ldg(reinterpret_cast<const float2*>(dxn + (size_t)r1 * C + col));
        *reinterpret_cast<float2*>(dx + (size_t)r1 * C + col) = make_float2(a.x + o[i][2], a.y + o[i][3]);
      }
    }
  }
}

// =========================================================================================
// backward, weight gradients (reduction over time on the tensor cores):
//   dWcur  = x[m]^T   . dpre[m]        -> filter[1], gate[1]
//   dWpast = x[m-d]^T . dpre[m]        -> filter[0], gate[0]
//   dWd    = z[m]^T   . dxn[m]         -> dense
//   dprebias[b] = sum_t dpre ; ddense_bias = sum dxn
// Each CTA strides over 64-row tiles, accumulates in registers and finishes with atomics.
// =========================================================================================
template <int C>
__global__ void __launch_bounds__(256, 2)
block_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dpre,
                   const float* __restrict__ zc, int ldz, const float* __restrict__ dxn,
                   float* __restrict__ gwf, float* __restrict__ gwg, float* __restrict__ gdense,
                   float* __restrict__ gprebias, float* __restrict__ gdense_bias, int M, int T, int d,
                   int is_last) {
  constexpr int TR = 64;              // rows per tile (K of the reduction)
  constexpr int PX = C + 8;           // pitches (floats), == 8 (mod 32) or conflict-free equivalent
  constexpr int PD = 2 * C + 8;
  constexpr int NT_CAT = (C / 16) * (2 * C / 8);   // 16x8 output tiles of one dWcat half
  constexpr int NT_D = (C / 16) * (C / 8);
  constexpr int NTILES = 2 * NT_CAT + NT_D;
  constexpr int PER_WARP = (NTILES + 7) / 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xc_s = reinterpret_cast<float*>(smem_raw);
  float* xp_s = xc_s + TR * PX;
  float* z_s = xp_s + TR * PX;
  float* dn_s = z_s + TR * PX;
  float* dp_s = dn_s + TR * PX;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float acc[PER_WARP][4];
#pragma unroll
  for (int i = 0; i < PER_WARP; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float bsum = 0.f;   // running bias-gradient sum of this thread's column
  int bcur = -1;      // batch element the running sum belongs to

  const int n_tiles = (M + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int m0 = tile * TR;
    __syncthreads();   // previous tile's readers are done
    // --- stage the tile (16-byte cp.async, zero fill for invalid rows) ---
    for (int i = threadIdx.x; i < TR * (C / 4); i += blockDim.x) {
      const int r = i / (C / 4), c4 = (i % (C / 4)) * 4;
      const int m = m0 + r;
      const bool v = m < M;
      const int tt = v ? m % T : 0;
      const size_t ms = v ? (size_t)m : 0;
      cp_async16(xc_s + r * PX + c4, x + ms * C + c4, v);
      const bool vp = v && tt >= d;
      cp_async16(xp_s + r * PX + c4, x + (vp ? ms - d : 0) * C + c4, vp);
      if (!is_last) {
        cp_async16(z_s + r * PX + c4, zc + ms * ldz + c4, v);
        cp_async16(dn_s + r * PX + c4, dxn + ms * C + c4, v);
      }
    }
    for (int i = threadIdx.x; i < TR * (2 * C / 4); i += blockDim.x) {
      const int r = i / (2 * C / 4), c4 = (i % (2 * C / 4)) * 4;
      const int m = m0 + r;
      const bool v = m < M;
      cp_async16(dp_s + r * PD + c4, dpre + (v ? (size_t)m : 0) * (2 * C) + c4, v);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // --- tensor-core reductions ---
#pragma unroll
    for (int i = 0; i < PER_WARP; ++i) {
      const int q = warp + 8 * i;
      if (q < NTILES) {
        const float* As;
        const float* Bs;
        int pb, mt, nt;
        if (q < 2 * NT_CAT) {
          const int h = q / NT_CAT, qq = q % NT_CAT;
          As = h ? xp_s : xc_s;
          Bs = dp_s;
          pb = PD;
          mt = qq / (2 * C / 8);
          nt = qq % (2 * C / 8);
        } else {
          const int qq = q - 2 * NT_CAT;
          As = z_s;
          Bs = dn_s;
          pb = PX;
          mt = qq / (C / 8);
          nt = qq % (C / 8);
        }
        if (!(is_last && q >= 2 * NT_CAT)) {
#pragma unroll
          for (int ks = 0; ks < TR / 8; ++ks) {
            const float* ar = As + (8 * ks + t) * PX + 16 * mt + g;
            uint32_t a0 = f2tf32(ar[0]), a1 = f2tf32(ar[8]);
            uint32_t a2 = f2tf32(ar[4 * PX]), a3 = f2tf32(ar[4 * PX + 8]);
            const float* br = Bs + (8 * ks + t) * pb + 8 * nt + g;
            uint32_t b0 = f2tf32(br[0]), b1 = f2tf32(br[4 * pb]);
            mma_tf32(acc[i], a0, a1, a2, a3, b0, b1);
          }
        }
      }
    }
    // --- bias gradients: one column per thread, flushed whenever the batch element changes ---
    if (threadIdx.x < 3 * C) {
      const bool is_pre = threadIdx.x < 2 * C;
      if (is_pre || !is_last) {
        const float* col = is_pre ? dp_s + threadIdx.x : dn_s + (threadIdx.x - 2 * C);
        const int pitch = is_pre ? PD : PX;
        for (int r = 0; r < TR; ++r) {
          const int m = m0 + r;
          if (m >= M) break;
          const int b = is_pre ? m / T : 0;
          if (b != bcur) {
            if (bcur >= 0) atomicAdd(is_pre ? gprebias + (size_t)bcur * 2 * C + threadIdx.x
                                            : gdense_bias + (threadIdx.x - 2 * C), bsum);
            bsum = 0.f;
            bcur = b;
          }
          bsum += col[r * pitch];
        }
      }
    }
  }
  if (threadIdx.x < 3 * C && bcur >= 0) {
    const bool is_pre = threadIdx.x < 2 * C;
    if (is_pre) atomicAdd(gprebias + (size_t)bcur * 2 * C + threadIdx.x, bsum);
    else if (gdense_bias) atomicAdd(gdense_bias + (threadIdx.x - 2 * C), bsum);
  }
  // --- flush the register accumulators ---
#pragma unroll
  for (int i = 0; i < PER_WARP; ++i) {
    const int q = warp + 8 * i;
    if (q >= NTILES) continue;
    if (q < 2 * NT_CAT) {
      const int h = q / NT_CAT, qq = q % NT_CAT;
      const int tap = h ? 0 : 1;
      const int mt = qq / (2 * C / 8), nt = qq % (2 * C / 8);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = 16 * mt + g + ((e & 2) ? 8 : 0);
        const int n = 8 * nt + 2 * t + (e & 1);
        float* dst = (n < C) ? gwf + (tap * C + r) * C + n : gwg + (tap * C + r) * C + (n - C);
        atomicAdd(dst, acc[i][e]);
      }
    } else if (!is_last) {
      const int qq = q - 2 * NT_CAT;
      const int mt = qq / (C / 8), nt = qq % (C / 8);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = 16 * mt + g + ((e & 2) ? 8 : 0);
        const int n = 8 * nt + 2 * t + (e & 1);
        atomicAdd(gdense + r * C + n, acc[i][e]);
      }
    }
  }
}

// =========================================================================================
// host launchers
// =========================================================================================
template <int C>
static int launch_fwd(const float* x, float* xout, float* zc, int ldz, const float* wf, const float* wg,
                      const float* dense, const float* prebias, const float* dense_bias, int M, int T,
                      int d, int is_last, cudaStream_t st) {
  using K = BlockCfg<C>;
  const size_t smem = 2 * sizeof(float2) * (C * K::NTP + (C / 2) * K::NDP);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(block_fwd3_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int n_tiles = (M + 15) / 16;
  int grid = (n_tiles + 7) / 8;
  const int cap = 2 * sm_count();
  if (grid > cap) grid = cap;
  block_fwd3_kernel<C><<<grid, 256, smem, st>>>(x, xout, zc, ldz, wf, wg, dense, prebias, dense_bias, M, T,
                                                d, is_last);
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_FWD);
  return 0;
}

template <int C>
static int launch_bwd(const float* x, const float* dxn, const float* dzs, int ldz, float* dx, float* dpre,
                      const float* zc, const float* wf, const float* wg, const float* dense,
                      const float* prebias, float* gwf, float* gwg, float* gdense, float* gprebias,
                      float* gdense_bias, int M, int T, int d, int is_last, cudaStream_t st) {
  using K = BlockCfg<C>;
  {
    const size_t smem = sizeof(float2) * (C * K::NTP + (C / 2) * K::DTP + 2 * C * K::NDP);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(block_bwd_dx_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr = true;
    }
    const int n_tiles = (M + 15) / 16;
    int grid = (n_tiles + 7) / 8;
    const int cap = sm_count();
    if (grid > cap) grid = cap;
    block_bwd_dx_kernel<C><<<grid, 256, smem, st>>>(x, dxn, dzs, ldz, dx, dpre, wf, wg, dense, prebias, M,
                                                    T, d, is_last);
    WN_CHECK_LAUNCH();
    prof_mark(st, PT_BLOCK_BWD_DX);
  }
  {
    const size_t smem = sizeof(float) * 64 * (4 * (C + 8) + (2 * C + 8));
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(block_wgrad_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr = true;
    }
    const int n_tiles = (M + 63) / 64;
    int grid = n_tiles;
    const int cap = 2 * sm_count();
    if (grid > cap) grid = cap;
    block_wgrad_kernel<C><<<grid, 256, smem, st>>>(x, dpre, zc, ldz, dxn, gwf, gwg, gdense, gprebias,
                                                   gdense_bias, M, T, d, is_last);
    WN_CHECK_LAUNCH();
    prof_mark(st, PT_BLOCK_WGRAD);
  }
  return 0;
}

static int g_block_impl = -1;   // -1 unset, 0 tcgen05, 1 mma.sync
void set_block_impl(int mma) { g_block_impl = mma ? 1 : 0; }
static bool use_mma_blocks() {
  if (g_block_impl < 0) {
    const char* e = getenv("WN_BLOCK_IMPL");
    g_block_impl = (e && strcmp(e, "mma") == 0) ? 1 : 0;
  }
  return g_block_impl == 1;
}

bool block_umma_enabled() { return !use_mma_blocks(); }

int block_fwd(const float* x, float* xout, float* zc, int ldz, const unsigned char* img,
              const float* wf, const float* wg, const float* dense, const float* prebias, const float* dense_bias,
              int M, int T, int d, int C, int is_last, cudaStream_t st) {
  if (C == 32 && !use_mma_blocks())
    return block_fwd_umma(x, xout, zc, ldz, img, wf, wg, dense, prebias, dense_bias, M / T, T, d,
                          is_last, st);
  int rc = -2;
  if (C == 32) rc = launch_fwd<32>(x, xout, zc, ldz, wf, wg, dense, prebias, dense_bias, M, T, d, is_last, st);
  if (C == 16) rc = launch_fwd<16>(x, xout, zc, ldz, wf, wg, dense, prebias, dense_bias, M, T, d, is_last, st);
  return rc;
}

int block_bwd(const float* x, const float* dxn, const float* dzs, int ldz, float* dx, float* dpre,
              const float* zc, const float* wf, const float* wg, const float* dense, const float* prebias,
              float* gwf, float* gwg, float* gdense, float* gprebias, float* gdense_bias, int M, int T,
              int d, int C, int is_last, cudaStream_t st) {
  if (C == 32)
    return launch_bwd<32>(x, dxn, dzs, ldz, dx, dpre, zc, wf, wg, dense, prebias, gwf, gwg, gdense, gprebias,
                          gdense_bias, M, T, d, is_last, st);
  if (C == 16)
    return launch_bwd<16>(x, dxn, dzs, ldz, dx, dpre, zc, wf, wg, dense, prebias, gwf, gwg, gdense, gprebias,
                          gdense_bias, M, T, d, is_last, st);
  return -2;
}

}  // namespace wn
