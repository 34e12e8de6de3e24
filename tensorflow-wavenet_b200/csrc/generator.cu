// Fast (incremental) generation as ONE persistent kernel per call: the sample loop, the
// per-layer delay lines (reference: one tf.FIFOQueue per layer, model.py:444-491), the
// per-step matvecs (model.py:332-387), post-processing (model.py:493-516), the float64
// softmax (model.py:619-626), temperature scaling + inverse-cdf draw (generate.py:228-241)
// and the feedback of the drawn sample all stay on the device; the host is not involved
// between samples (the reference does one sess.run per sample, generate.py:226).
//
// Work decomposition (v1): independent streams are batched SPB per CTA; the CTA walks the
// layer chain with the weights of layer l+1 prefetched (cp.async) into shared memory while
// layer l computes, every weight element fetched once per step is reused for all SPB streams,
// matvecs are k-split over thread groups and reduced through shared memory.
#include <cstdlib>
#include <cstring>

#include "../../include/wavenet_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "gen_common.h"

namespace wn {


// out[s][n] (+)= sum_k in[s][k] * W[k][n], W row-major [K][N] in GLOBAL memory, coalesced over n.
// Threads: column group cg = tid % NC4 (4 columns), k-slice kg = tid / NC4.  Result is left as
// k-slice partials in `part[kg][s][n]`; the caller reduces.
template <int SPB>
__device__ __forceinline__ int gemv_global(const float* __restrict__ W, int K, int N, const float* in, int in_ld,
                                           float* part) {
  const int nc4 = N >> 2;
  int kgroups = blockDim.x / nc4;
  if (kgroups < 1) kgroups = 1;
  if (kgroups > 32) kgroups = 32;
  const int kper = (K + kgroups - 1) / kgroups;
  for (int item = threadIdx.x; item < nc4 * kgroups; item += blockDim.x) {
    const int cg = item % nc4, kg = item / nc4;
    const int k0 = kg * kper;
    int k1 = k0 + kper;
    if (k1 > K) k1 = K;
    float4 acc[SPB];
#pragma unroll
    for (int s = 0; s < SPB; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* wp = reinterpret_cast<const float4*>(W) + cg;
    int k = k0;
    for (; k + 8 <= k1; k += 8) {
      float4 w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = __ldg(wp + (size_t)(k + u) * nc4);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int s = 0; s < SPB; ++s) {
          const float a = in[s * in_ld + k + u];
          acc[s].x = fmaf(a, w[u].x, acc[s].x);
          acc[s].y = fmaf(a, w[u].y, acc[s].y);
          acc[s].z = fmaf(a, w[u].z, acc[s].z);
          acc[s].w = fmaf(a, w[u].w, acc[s].w);
        }
      }
    }
    for (; k < k1; ++k) {
      const float4 w = __ldg(wp + (size_t)k * nc4);
#pragma unroll
      for (int s = 0; s < SPB; ++s) {
        const float a = in[s * in_ld + k];
        acc[s].x = fmaf(a, w.x, acc[s].x);
        acc[s].y = fmaf(a, w.y, acc[s].y);
        acc[s].z = fmaf(a, w.z, acc[s].z);
        acc[s].w = fmaf(a, w.w, acc[s].w);
      }
    }
#pragma unroll
    for (int s = 0; s < SPB; ++s)
      *reinterpret_cast<float4*>(part + ((size_t)kg * SPB + s) * N + cg * 4) = acc[s];
  }
  return kgroups;
}

template <int SPB>
__device__ __forceinline__ void reduce_parts(const float* part, int kgroups, int N, const float* __restrict__ bias,
                                             bool relu, float* out, int out_ld) {
  for (int i = threadIdx.x; i < SPB * N; i += blockDim.x) {
    const int s = i / N, n = i % N;
    float v = bias ? __ldg(bias + n) : 0.f;
    for (int g = 0; g < kgroups; ++g) v += part[((size_t)g * SPB + s) * N + n];
    out[s * out_ld + n] = relu ? fmaxf(v, 0.f) : v;
  }
}

template <int SPB, int C>
__global__ void __launch_bounds__(256, 1) generator_kernel(GenArgs a) {
  constexpr int NO = 2 * C;            // [f|g] outputs == [past|cur] inputs
  constexpr int G1 = 256 / NO;         // k-groups of the gated matvec
  constexpr int KP1 = NO / G1;
  constexpr int G2 = 256 / C;          // k-groups of the dense matvec
  constexpr int KP2 = C / G2;
  constexpr int WL = NO * NO + C * C;  // floats of one layer's staged weights
  extern __shared__ __align__(16) float sm[];
  const int LD = a.L * C;
  const int maxn = a.S > a.Q ? a.S : a.Q;
  float* wbuf = sm;                            // [2][WL]
  float* xin = wbuf + 2 * WL;                  // [SPB][NO]  = [past | cur]
  float* zs = xin + SPB * NO;                  // [SPB][C]
  float* zcat = zs + SPB * C;                  // [SPB][LD]
  float* v0 = zcat + SPB * LD;                 // [SPB][maxn]
  float* v1 = v0 + SPB * maxn;                 // [SPB][maxn]
  float* part = v1 + SPB * maxn;               // k-slice partials: PARTF floats per stream
  constexpr int PARTF = (16 * NO + 256 > 1024) ? 16 * NO + 256 : 1024;
  double* cdf = reinterpret_cast<double*>(part + SPB * PARTF);        // [SPB][Q]
  float* prebias = reinterpret_cast<float*>(cdf + SPB * a.Q);          // [SPB][L][NO]
  __shared__ int cur_id[SPB], prev_id[SPB], step0[SPB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s_base = blockIdx.x * SPB;

  auto stage_layer = [&](int l, int buf) {
    float* w = wbuf + buf * WL;
    const float* f = a.filter + (size_t)l * 2 * C * C;
    const float* g = a.gate + (size_t)l * 2 * C * C;
    const float* d = a.dense + (size_t)l * C * C;
    // Wcat[k][n]: k = tap*C + r, n<C filter, n>=C gate
    for (int i = tid; i < 2 * C * (C / 4); i += 256) {
      const int row = i / (C / 4), c4 = (i % (C / 4)) * 4;
      cp_async16(w + row * NO + c4, f + row * C + c4, true);
      cp_async16(w + row * NO + C + c4, g + row * C + c4, true);
    }
    for (int i = tid; i < C * (C / 4); i += 256) cp_async16(w + NO * NO + i * 4, d + i * 4, true);
    cp_async_commit();
  };

  // ---- per-launch setup: header, conditioning bias ----
  if (tid < SPB) {
    const int s = s_base + tid;
    if (s < a.streams) {
      prev_id[tid] = a.hdr[s * 4 + 0];
      step0[tid] = a.hdr[s * 4 + 1];
    } else {
      prev_id[tid] = -1;
      step0[tid] = 0;
    }
  }
  for (int i = tid; i < SPB * a.L * NO; i += 256) {
    const int n = i % NO, l = (i / NO) % a.L, s = i / (NO * a.L);
    const bool isg = n >= C;
    const int dd = isg ? n - C : n;
    float v = 0.f;
    if (a.use_biases) v = (isg ? a.gate_bias : a.filter_bias)[l * C + dd];
    if (a.G > 0 && a.gc_ids && s_base + s < a.streams && a.gc_ids[s_base + s] >= 0 && a.gc_ids[s_base + s] < a.gc_card) {
      const float* e = a.gc_embedding + (size_t)a.gc_ids[s_base + s] * a.G;
      const float* w = (isg ? a.gc_gate : a.gc_filter) + (size_t)l * a.G * C;
      for (int k = 0; k < a.G; ++k) v = fmaf(e[k], w[k * C + dd], v);
    }
    prebias[i] = v;
  }
  __syncthreads();

  for (int step = 0; step < a.n_steps; ++step) {
    // ---- input ids of this step ----
    if (tid < SPB) {
      const int s = s_base + tid;
      int id = -1;
      if (s < a.streams) {
        if (a.forced) id = a.forced[(size_t)s * a.n_steps + step];
        else if (step == 0) id = a.inputs[s];
        else id = cur_id[tid];   // sample drawn by the previous step (written below)
      }
      cur_id[tid] = id;
    }
    stage_layer(0, 0);
    __syncthreads();
    // ---- causal layer: x = Wc[0][prev] + Wc[1][cur]   (model.py:341-346) ----
    for (int i = tid; i < SPB * C; i += 256) {
      const int s = i / C, r = i % C;
      float v = 0.f;
      const int p = prev_id[s], c = cur_id[s];
      if (p >= 0 && p < a.Q) v += __ldg(a.causal + (size_t)p * C + r);
      if (c >= 0 && c < a.Q) v += __ldg(a.causal + (size_t)(a.Q + c) * C + r);
      xin[s * NO + C + r] = v;
    }
    // ---- dilated stack ----
    for (int l = 0; l < a.L; ++l) {
      const int buf = l & 1;
      const int d = a.dil[l];
      // past operand: ring slot (step index mod d) of every stream
      for (int i = tid; i < SPB * C; i += 256) {
        const int s = i / C, r = i % C;
        float v = 0.f;
        if (s_base + s < a.streams) {
          const int slot = (step0[s] + step) % d;
          v = a.rings[((size_t)(s_base + s) * a.sum_d + a.ring_off[l] + slot) * C + r];
        }
        xin[s * NO + r] = v;
      }
      if (l + 1 < a.L) stage_layer(l + 1, buf ^ 1);
      else cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      const float* w = wbuf + buf * WL;
      {  // gated matvec, k-split
        const int kg = tid / NO, n = tid % NO;
        float acc[SPB];
#pragma unroll
        for (int s = 0; s < SPB; ++s) acc[s] = 0.f;
#pragma unroll
        for (int k = 0; k < KP1; ++k) {
          const float wv = w[(kg * KP1 + k) * NO + n];
#pragma unroll
          for (int s = 0; s < SPB; ++s) acc[s] = fmaf(xin[s * NO + kg * KP1 + k], wv, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < SPB; ++s) part[(kg * SPB + s) * NO + n] = acc[s];
      }
      __syncthreads();
      for (int i = tid; i < SPB * C; i += 256) {
        const int s = i / C, dd = i % C;
        float f = prebias[(s * a.L + l) * NO + dd], g = prebias[(s * a.L + l) * NO + C + dd];
#pragma unroll
        for (int kg = 0; kg < G1; ++kg) {
          f += part[(kg * SPB + s) * NO + dd];
          g += part[(kg * SPB + s) * NO + C + dd];
        }
        const float z = tanhf(f) * (1.0f / (1.0f + expf(-g)));
        zs[s * C + dd] = z;
        zcat[s * LD + l * C + dd] = z;
      }
      __syncthreads();
      {  // dense matvec, k-split:  x += z . Wd + bd   (computed on the last layer too, model.py:377-380)
        const int kg = tid / C, n = tid % C;
        const float* wd = w + NO * NO;
        float acc[SPB];
#pragma unroll
        for (int s = 0; s < SPB; ++s) acc[s] = 0.f;
#pragma unroll
        for (int k = 0; k < KP2; ++k) {
          const float wv = wd[(kg * KP2 + k) * C + n];
#pragma unroll
          for (int s = 0; s < SPB; ++s) acc[s] = fmaf(zs[s * C + kg * KP2 + k], wv, acc[s]);
        }
        // stage partials after the gated partials have been consumed (same buffer, different region)
#pragma unroll
        for (int s = 0; s < SPB; ++s) part[16 * SPB * NO + (kg * SPB + s) * C + n] = acc[s];
      }
      __syncthreads();
      for (int i = tid; i < SPB * C; i += 256) {
        const int s = i / C, r = i % C;
        const float xold = xin[s * NO + C + r];
        float v = xold;
        if (a.use_biases) v += __ldg(a.dense_bias + l * C + r);
#pragma unroll
        for (int kg = 0; kg < G2; ++kg) v += part[16 * SPB * NO + (kg * SPB + s) * C + r];
        if (s_base + s < a.streams) {
          if (a.commit) {   // push_ops: enqueue the layer input
            const int slot = (step0[s] + step) % d;
            a.rings[((size_t)(s_base + s) * a.sum_d + a.ring_off[l] + slot) * C + r] = xold;
          } else {          // forward without push_ops: park it so that wn_gen_commit can enqueue later
            a.pending[((size_t)(s_base + s) * a.L + l) * C + r] = xold;
          }
        }
        xin[s * NO + C + r] = v;
      }
      __syncthreads();
    }
    cp_async_wait<0>();
    // ---- post-processing: sum of skips -> relu -> W1 -> relu -> W2   (model.py:505-514) ----
    {
      int kgs = gemv_global<SPB>(a.skip, LD, a.S, zcat, LD, part);
      __syncthreads();
      // bias of the skip sum = sum over layers of skip_bias
      for (int i = tid; i < SPB * a.S; i += 256) {
        const int s = i / a.S, n = i % a.S;
        float v = 0.f;
        if (a.use_biases)
          for (int l = 0; l < a.L; ++l) v += __ldg(a.skip_bias + (size_t)l * a.S + n);
        for (int g = 0; g < kgs; ++g) v += part[((size_t)g * SPB + s) * a.S + n];
        v0[s * maxn + n] = fmaxf(v, 0.f);
      }
      __syncthreads();
      kgs = gemv_global<SPB>(a.post1, a.S, a.S, v0, maxn, part);
      __syncthreads();
      reduce_parts<SPB>(part, kgs, a.S, a.use_biases ? a.post1_bias : nullptr, true, v1, maxn);
      __syncthreads();
      kgs = gemv_global<SPB>(a.post2, a.S, a.Q, v1, maxn, part);
      __syncthreads();
      reduce_parts<SPB>(part, kgs, a.Q, a.use_biases ? a.post2_bias : nullptr, false, v0, maxn);
      __syncthreads();
    }
    // ---- float64 softmax, temperature, inverse-cdf draw: one warp per stream ----
    if (warp < SPB) {
      const int s = warp;
      const float* lg = v0 + s * maxn;
      double mx = -1e300;
      for (int i = lane; i < a.Q; i += 32) mx = fmax(mx, (double)lg[i]);
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      double sum = 0.0;
      for (int i = lane; i < a.Q; i += 32) sum += exp((double)lg[i] - mx);
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      float* p = v1 + s * maxn;   // float32 probabilities (model.py:620-621 casts back)
      for (int i = lane; i < a.Q; i += 32) p[i] = (float)(exp((double)lg[i] - mx) / sum);
      __syncwarp();
      if (a.temperature != 1.0f) {   // generate.py:229-233, float32, parallel log-sum-exp
        float m2 = -INFINITY;
        for (int i = lane; i < a.Q; i += 32) {
          const float q = logf(p[i]) / a.temperature;
          p[i] = q;
          m2 = fmaxf(m2, q);
        }
        m2 = warp_max(m2);
        float s2 = 0.f;
        for (int i = lane; i < a.Q; i += 32) s2 += expf(p[i] - m2);
        s2 = warp_sum(s2);
        const float lse = m2 + logf(s2);
        for (int i = lane; i < a.Q; i += 32) p[i] = expf(p[i] - lse);
        __syncwarp();
      }
      const bool live = s_base + s < a.streams;
      if (a.proba_out && step == a.n_steps - 1 && live)
        for (int i = lane; i < a.Q; i += 32) a.proba_out[(size_t)(s_base + s) * a.Q + i] = p[i];
      int drawn = cur_id[s];
      if (a.uniforms && live) {
        double* c = cdf + s * a.Q;
        if (lane == 0) {   // np.cumsum order (sequential float64 adds)
          double run = 0.0;
          for (int i = 0; i < a.Q; ++i) { run += (double)p[i]; c[i] = run; }
        }
        __syncwarp();
        const double total = c[a.Q - 1];
        const double u = a.uniforms[(size_t)(s_base + s) * a.n_steps + step];
        int cnt = 0;
        for (int i = lane; i < a.Q; i += 32) cnt += (c[i] / total <= u) ? 1 : 0;   // searchsorted(..., 'right')
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        drawn = cnt < a.Q ? cnt : a.Q - 1;
        if (lane == 0) a.samples_out[(size_t)(s_base + s) * a.n_steps + step] = drawn;
      }
      if (lane == 0) {
        if (a.commit) prev_id[s] = cur_id[s];   // push of the causal queue (capacity 1)
        cur_id[s] = drawn;
      }
    }
    __syncthreads();
  }
  if (tid < SPB && s_base + tid < a.streams) {
    if (a.commit) {
      a.hdr[(s_base + tid) * 4 + 0] = prev_id[tid];
      a.hdr[(s_base + tid) * 4 + 1] = step0[tid] + a.n_steps;
      a.hdr[(s_base + tid) * 4 + 3] = 0;
    } else {
      a.hdr[(s_base + tid) * 4 + 2] = a.forced ? a.forced[(size_t)(s_base + tid) * a.n_steps + a.n_steps - 1]
                                               : (a.n_steps == 1 ? a.inputs[s_base + tid] : -1);
      a.hdr[(s_base + tid) * 4 + 3] = 1;
    }
  }
}

// enqueue the parked layer inputs of the last uncommitted single step (the reference's push_ops
// fetched after the fact)
__global__ void gen_commit_kernel(int32_t* hdr, const float* __restrict__ pending, float* __restrict__ rings,
                                  int streams, int L, int C, int sum_d, const int* __restrict__ dil_off) {
  const int s = blockIdx.x;
  if (s >= streams || hdr[s * 4 + 3] == 0) return;
  const int step = hdr[s * 4 + 1];
  for (int i = threadIdx.x; i < L * C; i += blockDim.x) {
    const int l = i / C, r = i % C;
    const int d = dil_off[2 * l], off = dil_off[2 * l + 1];
    rings[((size_t)s * sum_d + off + step % d) * C + r] = pending[((size_t)s * L + l) * C + r];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    hdr[s * 4 + 0] = hdr[s * 4 + 2];
    hdr[s * 4 + 1] = step + 1;
    hdr[s * 4 + 3] = 0;
  }
}

__global__ void gen_reset_kernel(int32_t* hdr, int streams) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < streams) {
    hdr[i * 4 + 0] = -1;
    hdr[i * 4 + 1] = 0;
    hdr[i * 4 + 2] = 0;
    hdr[i * 4 + 3] = 0;
  }
}

// standalone sampler: one warp per row
__global__ void sample_kernel(const float* __restrict__ proba, const double* __restrict__ uniforms, int rows, int Q,
                              int32_t* __restrict__ out) {
  extern __shared__ double cdf_s[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  double* c = cdf_s + (size_t)warp * Q;
  const float* p = proba + (size_t)row * Q;
  if (lane == 0) {
    double run = 0.0;
    for (int i = 0; i < Q; ++i) { run += (double)p[i]; c[i] = run; }
  }
  __syncwarp();
  const double total = c[Q - 1], u = uniforms[row];
  int cnt = 0;
  for (int i = lane; i < Q; i += 32) cnt += (c[i] / total <= u) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) out[row] = cnt;
}

static int sum_dil(const wn_config* c) {
  int s = 0;
  for (int i = 0; i < c->n_layers; ++i) s += c->dilations[i];
  return s;
}

static int64_t gen_hdr_bytes(const wn_config* c, int streams) {
  // header + (dilation, ring offset) table used by the commit kernel
  return ((int64_t)streams * 16 + (int64_t)c->n_layers * 8 + 255) / 256 * 256;
}
static int64_t gen_pending_bytes(const wn_config* c, int streams) {
  return ((int64_t)streams * c->n_layers * c->residual_channels * 4 + 255) / 256 * 256;
}
static int64_t gen_rings_bytes(const wn_config* c, int streams) {
  return ((int64_t)streams * sum_dil(c) * c->residual_channels * (int64_t)sizeof(float) + 255) / 256 * 256;
}

template <int SPB, int C>
static int launch_gen(const GenArgs& a, cudaStream_t st) {
  const int NO = 2 * C, LD = a.L * C;
  const int maxn = a.S > a.Q ? a.S : a.Q;
  const int partf = (16 * NO + 256 > 1024) ? 16 * NO + 256 : 1024;
  if (a.S > 1024 || a.Q > 1024) return -2;
  size_t fl = 2 * (NO * NO + C * C) + SPB * NO + SPB * C + (size_t)SPB * LD + 2 * SPB * maxn + (size_t)SPB * partf;
  size_t bytes = fl * sizeof(float);
  bytes = (bytes + 7) / 8 * 8;
  // the kernel carves cdf right after `part` -- keep the float count even so it is 8-byte aligned
  if (fl & 1) return -6;
  bytes += sizeof(double) * SPB * a.Q + sizeof(float) * SPB * a.L * NO;
  if (bytes > 227 * 1024) return -7;
  cudaError_t e = cudaFuncSetAttribute(generator_kernel<SPB, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  const int grid = (a.streams + SPB - 1) / SPB;
  generator_kernel<SPB, C><<<grid, 256, bytes, st>>>(a);
  WN_CHECK_LAUNCH();
  return 0;
}

}  // namespace wn

using namespace wn;

static int g_gen_lat_enabled = 1;

extern "C" {

int wn_debug_set_gen_impl(int32_t latency_kernel) {
  g_gen_lat_enabled = latency_kernel ? 1 : 0;
  return 0;
}

int64_t wn_gen_state_bytes(const wn_config* cfg, int32_t streams) {
  wn_layout lo;
  if (wn_param_layout(cfg, &lo) || streams < 1) return -1;
  // both generator kernels index filter / gate as [2C][C] and dense as [C][C]: one channel width (the training path
  // takes R != D through block_generic.cu, fast generation does not)
  if (cfg->dilation_channels != cfg->residual_channels) return -2;
  return gen_hdr_bytes(cfg, streams) + gen_pending_bytes(cfg, streams) + gen_rings_bytes(cfg, streams) +
         gen_lat_comm_bytes(cfg);      // tagged-word scratch of the latency-mode kernel (generator_lat.cu)
}

int wn_gen_reset(const wn_config* cfg, void* state, int32_t streams, wn_stream_t stream) {
  const int64_t bytes = wn_gen_state_bytes(cfg, streams);
  if (bytes < 0 || !state) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(state, 0, (size_t)bytes, st);
  if (e != cudaSuccess) return (int)e;
  gen_reset_kernel<<<(streams + 127) / 128, 128, 0, st>>>((int32_t*)state, streams);
  WN_CHECK_LAUNCH();
  // (dilation, ring offset) table right after the per-stream headers
  int tab[2 * WN_MAX_LAYERS];
  int off = 0;
  for (int i = 0; i < cfg->n_layers; ++i) { tab[2 * i] = cfg->dilations[i]; tab[2 * i + 1] = off; off += cfg->dilations[i]; }
  e = cudaMemcpyAsync((char*)state + (int64_t)streams * 16, tab, sizeof(int) * 2 * cfg->n_layers,
                      cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize(st);   // `tab` lives on this stack frame
  return (int)e;
}

int wn_gen_commit(const wn_config* cfg, void* state, int32_t streams, wn_stream_t stream) {
  if (wn_gen_state_bytes(cfg, streams) < 0 || !state) return -1;
  char* base = (char*)state;
  gen_commit_kernel<<<streams, 128, 0, (cudaStream_t)stream>>>(
      (int32_t*)base, (const float*)(base + gen_hdr_bytes(cfg, streams)),
      (float*)(base + gen_hdr_bytes(cfg, streams) + gen_pending_bytes(cfg, streams)), streams, cfg->n_layers,
      cfg->residual_channels, sum_dil(cfg), (const int*)(base + (int64_t)streams * 16));
  WN_CHECK_LAUNCH();
  return 0;
}

int wn_gen_run(const wn_config* cfg, const float* params, void* state, int32_t streams, const int32_t* inputs,
               const int32_t* forced, const int32_t* gc_ids, const double* uniforms, int32_t n_steps,
               float temperature, int32_t commit, int32_t* samples_out, float* proba_out, wn_stream_t stream) {
  wn_layout lo;
  int rc = wn_param_layout(cfg, &lo);
  if (rc) return rc;
  if (cfg->dilation_channels != cfg->residual_channels) return -2;      // see wn_gen_state_bytes
  if (cfg->scalar_input) return -2;      // model.py:601-603: "Scalar input is not supported by fast generation"
  if (!params || !state || streams < 1 || n_steps < 1) return -1;
  if (!inputs && !forced) return -1;
  if (uniforms && !samples_out) return -1;
  if (cfg->gc_channels > 0 && gc_ids && lo.gc_embedding < 0) return -1;
  if (!(temperature > 0.f)) return -1;
  GenArgs a;
  memset(&a, 0, sizeof(a));
  a.L = cfg->n_layers; a.C = cfg->residual_channels; a.S = cfg->skip_channels; a.Q = cfg->quantization_channels;
  a.G = gc_ids ? cfg->gc_channels : 0;
  a.gc_card = cfg->gc_cardinality;
  a.use_biases = cfg->use_biases;
  a.sum_d = sum_dil(cfg);
  a.streams = streams; a.n_steps = n_steps; a.commit = commit; a.temperature = temperature;
  auto PP = [&](int64_t off) { return off >= 0 ? params + off : (const float*)nullptr; };
  a.causal = PP(lo.causal); a.filter = PP(lo.filter); a.gate = PP(lo.gate); a.dense = PP(lo.dense);
  a.skip = PP(lo.skip); a.gc_filter = PP(lo.gc_filter); a.gc_gate = PP(lo.gc_gate);
  a.filter_bias = PP(lo.filter_bias); a.gate_bias = PP(lo.gate_bias); a.dense_bias = PP(lo.dense_bias);
  a.skip_bias = PP(lo.skip_bias); a.post1 = PP(lo.post1); a.post2 = PP(lo.post2);
  a.post1_bias = PP(lo.post1_bias); a.post2_bias = PP(lo.post2_bias); a.gc_embedding = PP(lo.gc_embedding);
  a.hdr = (int32_t*)state;
  a.pending = (float*)((char*)state + gen_hdr_bytes(cfg, streams));
  a.rings = (float*)((char*)state + gen_hdr_bytes(cfg, streams) + gen_pending_bytes(cfg, streams));
  a.inputs = inputs; a.forced = forced; a.gc_ids = gc_ids; a.uniforms = uniforms;
  a.samples_out = samples_out; a.proba_out = proba_out;
  int off = 0;
  for (int i = 0; i < a.L; ++i) { a.dil[i] = cfg->dilations[i]; a.ring_off[i] = off; off += cfg->dilations[i]; }
  cudaStream_t st = (cudaStream_t)stream;
  {   // one stream: latency-mode kernel (the whole GPU works on the sample chain); WN_GEN_IMPL=v1 disables it
    static int use_lat = -1;
    if (use_lat < 0) {
      const char* e = getenv("WN_GEN_IMPL");
      use_lat = (e && strcmp(e, "v1") == 0) ? 0 : 1;
    }
    static uint32_t launch_seq = 0;
    void* comm = (char*)state + gen_hdr_bytes(cfg, streams) + gen_pending_bytes(cfg, streams) + gen_rings_bytes(cfg, streams);
    if (use_lat && g_gen_lat_enabled && gen_lat_eligible(a)) return gen_lat_run(a, comm, ++launch_seq, st);
    // 2 .. 32 streams: the same layer-per-warp chain, the streams pipelined through it
    if (use_lat && g_gen_lat_enabled && gen_pipe_eligible(a)) return gen_pipe_run(a, comm, ++launch_seq, st);
  }
  const int nsm = sm_count();
  int spb = 1;
  if (streams > nsm) spb = 2;
  if (streams > 2 * nsm) spb = 4;
  {   // experiment knob: streams per CTA (more streams per CTA = fewer re-reads of the weights from L2)
    static int forced_spb = -1;
    if (forced_spb < 0) {
      const char* e = getenv("WN_GEN_SPB");
      forced_spb = e ? atoi(e) : 0;
    }
    if (forced_spb == 1 || forced_spb == 2 || forced_spb == 4) spb = forced_spb;
  }
  if (a.C == 32) {
    if (spb == 1) return launch_gen<1, 32>(a, st);
    if (spb == 2) return launch_gen<2, 32>(a, st);
    return launch_gen<4, 32>(a, st);
  } else if (a.C == 16) {
    if (spb == 1) return launch_gen<1, 16>(a, st);
    if (spb == 2) return launch_gen<2, 16>(a, st);
    return launch_gen<4, 16>(a, st);
  }
  return -2;
}

int wn_sample(const float* proba, const double* uniforms, int32_t rows, int32_t q, int32_t* out, wn_stream_t stream) {
  if (!proba || !uniforms || !out || rows < 1 || q < 1 || q > 4096) return -1;
  const int wpc = 4;
  sample_kernel<<<(rows + wpc - 1) / wpc, wpc * 32, sizeof(double) * wpc * q, (cudaStream_t)stream>>>(
      proba, uniforms, rows, q, out);
  WN_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
