// HBM-bound kernels around the dilated stack: mu-law companding (integer, bit exact),
// one-hot front end as a row gather, softmax cross entropy (forward + TF-style backprop),
// conditioning / bias helpers and the TF-semantics optimizers.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace wn {

// =========================================================================================
// mu-law                                                        wavenet/ops.py:65-85
// =========================================================================================
// encode(x) for |x| <= 1 is the number of decision thresholds <= x.  The Q-1 float32
// thresholds are produced on the host by bisection over the float32 restatement of the
// reference formula, which makes the device result bit-identical to it by construction
// (device logf differs from the host log in the last ulp, so evaluating the formula on the
// device would flip ~1e-5 of the bin edges).  Outside [-1,1] (the reference's librosa input
// never is) the formula itself is evaluated.
__device__ __forceinline__ int mulaw_one(float x, const float* __restrict__ thr, int Q) {
  if (fabsf(x) <= 1.0f) {
    int lo = 0, n = Q - 1;   // upper_bound: count of thr[i] <= x
    while (n > 0) {
      int half = n >> 1;
      if (thr[lo + half] <= x) { lo += half + 1; n -= half + 1; } else { n = half; }
    }
    return lo;
  }
  const float mu = (float)(Q - 1);
  float mag = __fdiv_rn(logf(__fadd_rn(1.0f, __fmul_rn(mu, fabsf(x)))), logf(__fadd_rn(1.0f, mu)));
  float sgn = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  float s = __fmul_rn(sgn, mag);
  float v = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(s, 1.0f), 2.0f), mu), 0.5f);
  return (int)v;   // truncation toward zero, like tf.cast(float -> int32)
}

__global__ void mulaw_encode_kernel(const float* __restrict__ audio, int64_t n,
                                    const float* __restrict__ thresholds, int Q, int32_t* __restrict__ ids) {
  extern __shared__ float thr[];
  for (int i = threadIdx.x; i < Q - 1; i += blockDim.x) thr[i] = thresholds[i];
  __syncthreads();
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = __ldg(reinterpret_cast<const float4*>(audio) + i);
    int4 o;
    o.x = mulaw_one(a.x, thr, Q);
    o.y = mulaw_one(a.y, thr, Q);
    o.z = mulaw_one(a.z, thr, Q);
    o.w = mulaw_one(a.w, thr, Q);
    reinterpret_cast<int4*>(ids)[i] = o;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    ids[i] = mulaw_one(audio[i], thr, Q);
}

int mulaw_encode(const float* audio, int64_t n, const float* thresholds, int Q, int32_t* ids, cudaStream_t st) {
  if (n < 0 || Q < 2 || Q > 8192) return -1;
  if (n == 0) return 0;
  if (((uintptr_t)audio & 15) || ((uintptr_t)ids & 15)) return -4;
  int64_t blocks = ((n >> 2) + 255) / 256;
  const int cap = 8 * sm_count();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  mulaw_encode_kernel<<<(int)blocks, 256, sizeof(float) * (Q - 1), st>>>(audio, n, thresholds, Q, ids);
  WN_CHECK_LAUNCH();
  return 0;
}

// decode = gather from the Q-entry float32 table of the reference formula (ops.py:76-85);
// ids outside [0,Q) evaluate the formula directly.
__global__ void mulaw_decode_kernel(const int32_t* __restrict__ ids, int64_t n, const float* __restrict__ lut,
                                    int Q, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int id = ids[i];
    float r;
    if (id >= 0 && id < Q) {
      r = __ldg(lut + id);
    } else {
      const float mu = (float)(Q - 1);
      float s = __fadd_rn(__fmul_rn(2.0f, __fdiv_rn((float)id, mu)), -1.0f);
      float mag = __fmul_rn((float)(1.0 / (double)(Q - 1)), __fadd_rn(powf(1.0f + mu, fabsf(s)), -1.0f));
      r = (s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f)) * mag;
    }
    out[i] = r;
  }
}

int mulaw_decode(const int32_t* ids, int64_t n, const float* lut, int Q, float* out, cudaStream_t st) {
  if (n < 0 || Q < 2) return -1;
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  const int cap = 8 * sm_count();
  if (blocks > cap) blocks = cap;
  mulaw_decode_kernel<<<(int)blocks, 256, 0, st>>>(ids, n, lut, Q, out);
  WN_CHECK_LAUNCH();
  return 0;
}

// =========================================================================================
// front end: one_hot (model.py:518-531) + causal layer (model.py:227-234) as a row gather
//   x0[m] = Wc[0][id[m-1]] (t>0) + Wc[1][id[m]];   out-of-range ids are all-zero one-hot rows
// =========================================================================================
// xs (optional, R == 32): the same rows also as fp16 split rows [hi 32 | lo 32] (input of the first forward layer)
__global__ void frontend_fwd_kernel(const int32_t* __restrict__ ids, const float* __restrict__ wc,
                                    float* __restrict__ x0, int M, int T, int Q, int R, __half* __restrict__ xs) {
  const int r4 = R >> 2;
  const int64_t total = (int64_t)M * r4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int m = (int)(i / r4), c = (int)(i % r4) * 4;
    const int t = m % T;
    const int cur = __ldg(ids + m);
    const int prev = t > 0 ? __ldg(ids + m - 1) : -1;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prev >= 0 && prev < Q) {
      float4 a = __ldg(reinterpret_cast<const float4*>(wc + (size_t)prev * R + c));
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    if (cur >= 0 && cur < Q) {
      float4 a = __ldg(reinterpret_cast<const float4*>(wc + (size_t)(Q + cur) * R + c));
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    *reinterpret_cast<float4*>(x0 + (size_t)m * R + c) = o;
    if (xs) {
      const float v[4] = {o.x, o.y, o.z, o.w};
      __align__(8) __half h[4], l[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h[q] = __float2half_rn(v[q]);
        l[q] = __float2half_rn(v[q] - __half2float(h[q]));
      }
      __half* row = xs + (size_t)m * 64;
      *reinterpret_cast<uint2*>(row + c) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(row + 32 + c) = *reinterpret_cast<const uint2*>(l);
    }
  }
}
int frontend_fwd(const int32_t* ids, const float* wc, float* x0, int M, int T, int Q, int R, void* xs, cudaStream_t st) {
  if (xs && R != 32) return -1;
  if (M <= 0 || (R & 3)) return -1;
  int64_t blocks = ((int64_t)M * (R / 4) + 255) / 256;
  const int cap = 16 * sm_count();
  if (blocks > cap) blocks = cap;
  frontend_fwd_kernel<<<(int)blocks, 256, 0, st>>>(ids, wc, x0, M, T, Q, R, (__half*)xs);
  WN_CHECK_LAUNCH();
  return 0;
}

// gwc[0][id[m-1]] += dx0[m] (t>0) ; gwc[1][id[m]] += dx0[m]   (histogram-like scatter add)
template <bool SMEM>
__global__ void frontend_bwd_kernel(const int32_t* __restrict__ ids, const float* __restrict__ dx0,
                                    float* __restrict__ gwc, int M, int T, int Q, int R, int rows_per_cta) {
  extern __shared__ float acc[];
  const int n_acc = 2 * Q * R;
  if (SMEM) {
    for (int i = threadIdx.x; i < n_acc; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
  }
  float* dst = SMEM ? acc : gwc;
  const int m0 = blockIdx.x * rows_per_cta;
  int m1 = m0 + rows_per_cta;
  if (m1 > M) m1 = M;
  const int rows_per_it = blockDim.x / R;
  const int rr = threadIdx.x / R, c = threadIdx.x % R;
  if (rr < rows_per_it) {
    // four rows per thread per trip, all loads first: one row per trip left the loop bound by one global-load latency per
    // row (85 trips per CTA, ~60 us for a kernel that moves 13 MB)
    int m = m0 + rr;
    for (; m + 3 * rows_per_it < m1; m += 4 * rows_per_it) {
      float v[4];
      int cur[4], prev[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int mm = m + u * rows_per_it;
        v[u] = __ldg(dx0 + (size_t)mm * R + c);
        cur[u] = __ldg(ids + mm);
        prev[u] = (mm % T) > 0 ? __ldg(ids + mm - 1) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (prev[u] >= 0 && prev[u] < Q) atomicAdd(dst + (size_t)prev[u] * R + c, v[u]);
        if (cur[u] >= 0 && cur[u] < Q) atomicAdd(dst + (size_t)(Q + cur[u]) * R + c, v[u]);
      }
    }
    for (; m < m1; m += rows_per_it) {
      const float v = __ldg(dx0 + (size_t)m * R + c);
      const int t = m % T;
      const int cur = __ldg(ids + m);
      const int prev = t > 0 ? __ldg(ids + m - 1) : -1;
      if (prev >= 0 && prev < Q) atomicAdd(dst + (size_t)prev * R + c, v);
      if (cur >= 0 && cur < Q) atomicAdd(dst + (size_t)(Q + cur) * R + c, v);
    }
  }
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_acc; i += blockDim.x) {
      const float v = acc[i];
      if (v != 0.f) atomicAdd(gwc + i, v);
    }
  }
}
int frontend_bwd(const int32_t* ids, const float* dx0, float* gwc, int M, int T, int Q, int R, cudaStream_t st) {
  if (M <= 0 || R > 256 || (256 % R)) return -1;
  const size_t smem = sizeof(float) * 2 * Q * R;
  int ctas = sm_count();
  int rows = (M + ctas - 1) / ctas;
  if (rows < 64) rows = 64;
  ctas = (M + rows - 1) / rows;
  if (smem <= 160 * 1024) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(frontend_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr = true;
    }
    frontend_bwd_kernel<true><<<ctas, 256, smem, st>>>(ids, dx0, gwc, M, T, Q, R, rows);
  } else {
    frontend_bwd_kernel<false><<<ctas, 256, 0, st>>>(ids, dx0, gwc, M, T, Q, R, rows);
  }
  WN_CHECK_LAUNCH();
  return 0;
}

// =========================================================================================
// softmax cross entropy against the next sample               model.py:654-666
//   target(m) = id[m+1] if t < T-1 (and in range) else none (all-zero label row)
//   loss = mean over all B*T rows;  backprop = (softmax - onehot) * scale  (TF's xent kernel
//   emits softmax - labels even for the all-zero label row)
// One warp per row, float4 coalesced; logits are overwritten with the gradient.
// =========================================================================================
template <int NV>   // float4 vectors per lane: Q <= 128*NV
__global__ void __launch_bounds__(256)
xent_kernel(float* __restrict__ logits, const int32_t* __restrict__ ids, int M, int T, int Q, float scale,
            float* __restrict__ partials, int write_grad, __half* __restrict__ g16, float scale16) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  double local = 0.0;
  for (int m = blockIdx.x * wpc + warp; m < M; m += gridDim.x * wpc) {
    float* row = logits + (size_t)m * Q;
    float4 v[NV];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      if (c < Q) {
        v[i] = *reinterpret_cast<const float4*>(row + c);
        mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      }
    }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      if (c < Q) {
        v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx);
        v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    s = warp_sum(s);
    const int t = m % T;
    int target = -1;
    if (t < T - 1) {
      target = __ldg(ids + m + 1);
      if (target < 0 || target >= Q) target = -1;
    }
    const float inv = 1.0f / s;
    float pt = 0.f;   // probability of the target class (lane that owns it)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      if (c < Q) {
        float p0 = v[i].x * inv, p1 = v[i].y * inv, p2 = v[i].z * inv, p3 = v[i].w * inv;
        if (target >= c && target < c + 4) {
          const int k = target - c;
          pt = k == 0 ? p0 : (k == 1 ? p1 : (k == 2 ? p2 : p3));
          if (k == 0) p0 -= 1.f; else if (k == 1) p1 -= 1.f; else if (k == 2) p2 -= 1.f; else p3 -= 1.f;
        }
        if (write_grad) {
          float4 o;
          o.x = round_tf32(p0 * scale); o.y = round_tf32(p1 * scale);
          o.z = round_tf32(p2 * scale); o.w = round_tf32(p3 * scale);
          *reinterpret_cast<float4*>(row + c) = o;
        }
        {
          if (g16) {      // the same gradient in the scaled fp16 domain of the input-gradient GEMM chain
            __half2 h[2] = {__floats2half2_rn(p0 * scale16, p1 * scale16), __floats2half2_rn(p2 * scale16, p3 * scale16)};
            *reinterpret_cast<uint2*>(g16 + (size_t)m * Q + c) = *reinterpret_cast<uint2*>(h);
          }
        }
      }
    }
    pt = warp_sum(pt);   // exactly one lane contributed
    if (lane == 0 && target >= 0) local += (double)(-logf(fmaxf(pt, 1e-37f)));
  }
  __shared__ double red[8];
  if (lane == 0) red[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < wpc; ++i) s += red[i];
    partials[blockIdx.x] = (float)s;
  }
}

__global__ void xent_finalize_kernel(const float* __restrict__ partials, int n, float scale,
                                     float* __restrict__ loss_out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partials[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (float)(red[0] * (double)scale);
}

int xent_finalize(const float* partials, int n, float scale, float* loss_out, cudaStream_t st) {
  xent_finalize_kernel<<<1, 256, 0, st>>>(partials, n, scale, loss_out);
  WN_CHECK_LAUNCH();
  return 0;
}

int softmax_xent(float* logits, const int32_t* ids, int M, int T, int Q, float scale, float* partials,
                 int n_partials, float* loss_out, int write_grad, void* g16v, float scale16, cudaStream_t st) {
  __half* g16 = (__half*)g16v;
  if (M <= 0 || (Q & 3) || Q > 1024 || n_partials < 1) return -1;
  int grid = (M + 7) / 8;
  if (grid > n_partials) grid = n_partials;
  if (Q <= 128) xent_kernel<1><<<grid, 256, 0, st>>>(logits, ids, M, T, Q, scale, partials, write_grad, g16, scale16);
  else if (Q <= 256) xent_kernel<2><<<grid, 256, 0, st>>>(logits, ids, M, T, Q, scale, partials, write_grad, g16, scale16);
  else if (Q <= 512) xent_kernel<4><<<grid, 256, 0, st>>>(logits, ids, M, T, Q, scale, partials, write_grad, g16, scale16);
  else xent_kernel<8><<<grid, 256, 0, st>>>(logits, ids, M, T, Q, scale, partials, write_grad, g16, scale16);
  WN_CHECK_LAUNCH();
  xent_finalize_kernel<<<1, 256, 0, st>>>(partials, grid, scale, loss_out);
  WN_CHECK_LAUNCH();
  return 0;
}

// =========================================================================================
// conditioning / bias helpers                               model.py:272-290,533-562
// =========================================================================================
__global__ void cond_bias_fwd_kernel(float* __restrict__ prebias, const float* __restrict__ filter_bias,
                                     const float* __restrict__ gate_bias, const float* __restrict__ gc_filter,
                                     const float* __restrict__ gc_gate, const float* __restrict__ emb_table,
                                     const int32_t* __restrict__ gc_ids, int L, int B, int D, int G, int card) {
  const int total = L * B * 2 * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % (2 * D), b = (i / (2 * D)) % B, l = i / (2 * D * B);
    const bool is_g = n >= D;
    const int dd = is_g ? n - D : n;
    float v = 0.f;
    const float* bias = is_g ? gate_bias : filter_bias;
    if (bias) v = bias[l * D + dd];
    // an id outside [0, card) contributes a zero embedding (tf.nn.embedding_lookup on the GPU) and never reads past the table
    if (G > 0 && gc_ids[b] >= 0 && gc_ids[b] < card) {
      const float* w = (is_g ? gc_gate : gc_filter) + (size_t)l * G * D;
      const float* e = emb_table + (size_t)gc_ids[b] * G;
      for (int k = 0; k < G; ++k) v = fmaf(e[k], w[k * D + dd], v);
    }
    prebias[i] = v;
  }
}

int cond_bias_fwd(float* prebias, const float* filter_bias, const float* gate_bias, const float* gc_filter,
                  const float* gc_gate, const float* emb_table, const int32_t* gc_ids, int L, int B, int D,
                  int G, int card, cudaStream_t st) {
  const int total = L * B * 2 * D;
  cond_bias_fwd_kernel<<<(total + 255) / 256, 256, 0, st>>>(prebias, filter_bias, gate_bias, gc_filter, gc_gate,
                                                             emb_table, gc_ids, L, B, D, G, card);
  WN_CHECK_LAUNCH();
  return 0;
}

__global__ void cond_bias_bwd_kernel(const float* __restrict__ gpre, float* __restrict__ gfb,
                                     float* __restrict__ ggb, const float* __restrict__ gc_filter,
                                     const float* __restrict__ gc_gate, float* __restrict__ ggc_filter,
                                     float* __restrict__ ggc_gate, const float* __restrict__ emb_table,
                                     float* __restrict__ gemb_table, const int32_t* __restrict__ gc_ids, int L,
                                     int B, int D, int G, int card) {
  const int n_bias = L * 2 * D;
  const int n_w = L * G * 2 * D;
  const int n_e = B * G;
  const int total = n_bias + n_w + n_e;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < n_bias) {
      const int n = i % (2 * D), l = i / (2 * D);
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += gpre[((size_t)l * B + b) * 2 * D + n];
      if (n < D) { if (gfb) gfb[l * D + n] += s; }
      else { if (ggb) ggb[l * D + (n - D)] += s; }
    } else if (i < n_bias + n_w) {
      const int j = i - n_bias;
      const int n = j % (2 * D), k = (j / (2 * D)) % G, l = j / (2 * D * G);
      float s = 0.f;
      for (int b = 0; b < B; ++b)
        if (gc_ids[b] >= 0 && gc_ids[b] < card)      // out-of-range id: zero embedding row, no gradient
          s = fmaf(emb_table[(size_t)gc_ids[b] * G + k], gpre[((size_t)l * B + b) * 2 * D + n], s);
      if (n < D) ggc_filter[((size_t)l * G + k) * D + n] += s;
      else ggc_gate[((size_t)l * G + k) * D + (n - D)] += s;
    } else {
      const int j = i - n_bias - n_w;
      const int k = j % G, b = j / G;
      float s = 0.f;
      for (int l = 0; l < L; ++l) {
        const float* gp = gpre + ((size_t)l * B + b) * 2 * D;
        const float* wf = gc_filter + ((size_t)l * G + k) * D;
        const float* wg = gc_gate + ((size_t)l * G + k) * D;
        for (int dd = 0; dd < D; ++dd) s += gp[dd] * wf[dd] + gp[D + dd] * wg[dd];
      }
      if (gemb_table && gc_ids[b] >= 0 && gc_ids[b] < card) atomicAdd(gemb_table + (size_t)gc_ids[b] * G + k, s);
    }
  }
}

int cond_bias_bwd(const float* gprebias, float* gfilter_bias, float* ggate_bias, const float* gc_filter,
                  const float* gc_gate, float* ggc_filter, float* ggc_gate, const float* emb_table,
                  float* gemb_table, const int32_t* gc_ids, int L, int B, int D, int G, int card,
                  cudaStream_t st) {
  const int total = L * 2 * D + L * G * 2 * D + B * G;
  cond_bias_bwd_kernel<<<(total + 127) / 128, 128, 0, st>>>(gprebias, gfilter_bias, ggate_bias, gc_filter, gc_gate,
                                                             ggc_filter, ggc_gate, emb_table, gemb_table, gc_ids,
                                                             L, B, D, G, card);
  WN_CHECK_LAUNCH();
  return 0;
}

// float64 softmax of one row of logits, cast to float32 (model.py:584-585, 620-621)
__global__ void softmax_f64_kernel(const float* __restrict__ logits, int Q, float* __restrict__ proba) {
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double mx = -1e300;
  for (int i = tid; i < Q; i += blockDim.x) mx = fmax(mx, (double)logits[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < nw; ++w) mx = fmax(mx, red[w]);
  __syncthreads();
  double sum = 0.0;
  for (int i = tid; i < Q; i += blockDim.x) sum += exp((double)logits[i] - mx);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.0;
  for (int w = 0; w < nw; ++w) sum += red[w];
  for (int i = tid; i < Q; i += blockDim.x) proba[i] = (float)(exp((double)logits[i] - mx) / sum);
}
int softmax_f64(const float* logits, int Q, float* proba, cudaStream_t st) {
  softmax_f64_kernel<<<1, 256, 0, st>>>(logits, Q, proba);
  WN_CHECK_LAUNCH();
  return 0;
}

__global__ void add_l2_kernel(float* __restrict__ loss, const float* __restrict__ p, int64_t n, float coef) {
  __shared__ float red[8];
  float s = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s = fmaf(p[i], p[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(loss, 0.5f * coef * t);
  }
}
int add_l2(float* loss, const float* params, int64_t n, float coef, cudaStream_t st) {
  int64_t blocks = (n + 255) / 256;
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  add_l2_kernel<<<(int)blocks, 256, 0, st>>>(loss, params, n, coef);
  WN_CHECK_LAUNCH();
  return 0;
}

__global__ void skip_bias_sum_kernel(const float* __restrict__ skip_bias, int L, int S, float* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float v = 0.f;
  for (int l = 0; l < L; ++l) v += skip_bias[(size_t)l * S + s];
  out[s] = v;
}
int skip_bias_sum(const float* skip_bias, int L, int S, float* out, cudaStream_t st) {
  skip_bias_sum_kernel<<<(S + 127) / 128, 128, 0, st>>>(skip_bias, L, S, out);
  WN_CHECK_LAUNCH();
  return 0;
}

__global__ void bcast_rows_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int rows) {
  const int total = rows * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) dst[i] += src[i % n];
}
int bcast_rows(const float* src, int n, float* dst, int rows, cudaStream_t st) {
  const int total = rows * n;
  bcast_rows_kernel<<<(total + 255) / 256, 256, 0, st>>>(src, n, dst, rows);
  WN_CHECK_LAUNCH();
  return 0;
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n, int round_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = dst[i] + src[i];
    dst[i] = round_out ? round_tf32(v) : v;
  }
}
int add_inplace(float* dst, const float* src, int64_t n, int round_out, cudaStream_t st) {
  int64_t blocks = (n + 255) / 256;
  const int cap = 16 * sm_count();
  if (blocks > cap) blocks = cap;
  add_inplace_kernel<<<(int)blocks, 256, 0, st>>>(dst, src, n, round_out);
  WN_CHECK_LAUNCH();
  return 0;
}

// dst = dst + grad * (act > 0)
__global__ void relu_mask_add_kernel(float* __restrict__ dst, const float* __restrict__ grad,
                                     const float* __restrict__ act, int64_t n, int round_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = dst[i] + (act[i] > 0.f ? grad[i] : 0.f);
    dst[i] = round_out ? round_tf32(v) : v;
  }
}
int relu_mask_add(float* dst, const float* grad, const float* act, int64_t n, int round_out, cudaStream_t st) {
  int64_t blocks = (n + 255) / 256;
  const int cap = 16 * sm_count();
  if (blocks > cap) blocks = cap;
  relu_mask_add_kernel<<<(int)blocks, 256, 0, st>>>(dst, grad, act, n, round_out);
  WN_CHECK_LAUNCH();
  return 0;
}

// =========================================================================================
// optimizers with TensorFlow-0.10 update rules (wavenet/ops.py:6-24).  Written with
// explicitly rounded operations (no FMA contraction) so that they match a float32 NumPy
// evaluation of the same formulas bit for bit.
// =========================================================================================
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_t, float b1, float omb1, float b2,
                            float omb2, float eps, float l2, float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = g[i];
    if (gscale != 1.0f) gi = __fmul_rn(gi, gscale);
    if (l2 != 0.0f) gi = __fadd_rn(gi, __fmul_rn(l2, w[i]));
    const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(omb1, gi));
    const float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(__fmul_rn(omb2, gi), gi));
    m[i] = mi;
    v[i] = vi;
    w[i] = __fsub_rn(w[i], __fdiv_rn(__fmul_rn(lr_t, mi), __fadd_rn(__fsqrt_rn(vi), eps)));
  }
}
int optim_adam(float* w, const float* g, float* m, float* v, int64_t n, double lr_t, double beta1, double beta2,
               double eps, float l2, float gscale, cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  const int cap = 8 * sm_count();
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(int)blocks, 256, 0, st>>>(w, g, m, v, n, (float)lr_t, (float)beta1, (float)(1.0 - beta1), (float)beta2,
                                           (float)(1.0 - beta2), (float)eps, l2, gscale);
  WN_CHECK_LAUNCH();
  return 0;
}

__global__ void momentum_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ a,
                                int64_t n, float lr, float mu, float l2, float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = g[i];
    if (gscale != 1.0f) gi = __fmul_rn(gi, gscale);
    if (l2 != 0.0f) gi = __fadd_rn(gi, __fmul_rn(l2, w[i]));
    const float ai = __fadd_rn(__fmul_rn(mu, a[i]), gi);
    a[i] = ai;
    w[i] = __fsub_rn(w[i], __fmul_rn(lr, ai));
  }
}
int optim_momentum(float* w, const float* g, float* a, int64_t n, double lr, double mu, float l2, float gscale,
                   cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  const int cap = 8 * sm_count();
  if (blocks > cap) blocks = cap;
  momentum_kernel<<<(int)blocks, 256, 0, st>>>(w, g, a, n, (float)lr, (float)mu, l2, gscale);
  WN_CHECK_LAUNCH();
  return 0;
}

__global__ void rmsprop_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ ms,
                               float* __restrict__ mom, int64_t n, float lr, float decay, float omd, float mu,
                               float eps, float l2, float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = g[i];
    if (gscale != 1.0f) gi = __fmul_rn(gi, gscale);
    if (l2 != 0.0f) gi = __fadd_rn(gi, __fmul_rn(l2, w[i]));
    const float msi = __fadd_rn(__fmul_rn(decay, ms[i]), __fmul_rn(__fmul_rn(omd, gi), gi));
    const float mi = __fadd_rn(__fmul_rn(mu, mom[i]),
                               __fdiv_rn(__fmul_rn(lr, gi), __fsqrt_rn(__fadd_rn(msi, eps))));
    ms[i] = msi;
    mom[i] = mi;
    w[i] = __fsub_rn(w[i], mi);
  }
}
int optim_rmsprop(float* w, const float* g, float* ms, float* mom, int64_t n, double lr, double decay, double mu,
                  double eps, float l2, float gscale, cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  const int cap = 8 * sm_count();
  if (blocks > cap) blocks = cap;
  rmsprop_kernel<<<(int)blocks, 256, 0, st>>>(w, g, ms, mom, n, (float)lr, (float)decay, (float)(1.0 - decay), (float)mu,
                                              (float)eps, l2, gscale);
  WN_CHECK_LAUNCH();
  return 0;
}

// =========================================================================================
// generic dilated causal convolution (the public ops.causal_conv, wavenet/ops.py:46-62).
// Closed form of pad -> time_to_batch -> conv1d(SAME) -> batch_to_time -> slice for any
// filter width W:   y[t] = sum_k w[k] . x[t - (W-1 + (W-1)/2 - k) * d]   (zero before t=0;
// for W == 2 this is x[t-d].w[0] + x[t].w[1]; for W > 2 the SAME padding of this snapshot
// adds the extra (W-1)/2 lag).  Plain fp32 FMA -- exact for the integer-valued
// reference tests; the training path uses the fused tensor-core block kernels instead.
// =========================================================================================
__global__ void causal_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                   int M, int T, int cin, int cout, int width, int d) {
  const int64_t total = (int64_t)M * cout;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int left = (width - 1) / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int m = (int)(i / cout), co = (int)(i % cout);
    const int t = m % T;
    float acc = 0.f;
    for (int k = 0; k < width; ++k) {
      const int lag = (width - 1 + left - k) * d;
      if (t - lag < 0) continue;
      const float* xr = x + (size_t)(m - lag) * cin;
      const float* wk = w + (size_t)k * cin * cout + co;
      for (int ci = 0; ci < cin; ++ci) acc = fmaf(xr[ci], wk[(size_t)ci * cout], acc);
    }
    y[i] = acc;
  }
}
// scalar_input front end (model.py:143-153, 227-234), gradient of the [width, 1, R] filter: one block per chunk of time
// steps, thread = (tap, channel) pairs, partial sums added to the gradient with one atomic per pair and block
__global__ void scalar_frontend_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dx0, float* __restrict__ gw,
                                           int M, int T, int R, int width, int rows_per_block) {
  const int m0 = blockIdx.x * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  const int left = (width - 1) / 2;
  for (int pidx = threadIdx.x; pidx < width * R; pidx += blockDim.x) {
    const int k = pidx / R, r = pidx % R;
    const int lag = width - 1 + left - k;      // the tap's lag, dilation 1
    float acc = 0.f;
    for (int m = m0; m < m1; ++m) {
      const int t = m % T;
      if (t - lag >= 0) acc = fmaf(x[m - lag], dx0[(size_t)m * R + r], acc);
    }
    atomicAdd(gw + pidx, acc);
  }
}
int scalar_frontend_bwd(const float* x, const float* dx0, float* gw, int M, int T, int R, int width, cudaStream_t st) {
  if (M <= 0 || R < 1 || width < 1) return -1;
  const int rows = 256;
  scalar_frontend_bwd_kernel<<<(M + rows - 1) / rows, 256, 0, st>>>(x, dx0, gw, M, T, R, width, rows);
  WN_CHECK_LAUNCH();
  return 0;
}

int causal_conv(const float* x, const float* w, float* y, int M, int T, int cin, int cout, int width, int d,
                cudaStream_t st) {
  if (M <= 0 || cin < 1 || cout < 1 || width < 1 || d < 1) return -1;
  int64_t blocks = ((int64_t)M * cout + 255) / 256;
  const int cap = 16 * sm_count();
  if (blocks > cap) blocks = cap;
  causal_conv_kernel<<<(int)blocks, 256, 0, st>>>(x, w, y, M, T, cin, cout, width, d);
  WN_CHECK_LAUNCH();
  return 0;
}

}  // namespace wn
