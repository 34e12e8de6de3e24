// Fast generation, latency mode: ONE stream spread over the whole GPU by a persistent cooperative kernel.
// Reference: wavenet/model.py:332-387 (_generator_conv / _generator_dilation_layer), :444-516
// (_create_generator, one tf.FIFOQueue per layer), :592-626 (predict_proba_incremental) and the sampling
// loop of generate.py:213-241 (one sess.run per sample there; here the sample loop never leaves the device).
//
// A sample is a dependent chain of L+3 stages, so the design minimises the latency of each hop:
//   * chain CTAs   : LPC = 8 layers per CTA, ONE WARP PER LAYER with that layer's weights held in registers
//                    (2*32*64 filter/gate + 32*32 dense floats = 160 registers per lane, packed-FMA pairs).
//                    The half of the gated pre-activation that multiplies the delay-line output (x[t-d]) is
//                    computed while the warp waits for its input; a layer then costs one 32x64 and one
//                    32x32 register-resident mat-vec.  Warps hand the 32-float residual to the next layer
//                    through shared memory (next CTA: through L2) as 64-bit words {value, tag}: a word is valid
//                    when its tag names the current launch and step, so no fences or flags are needed.
//                    Delay lines stay where the reference's queues semantics put them (ring slot = t mod d)
//                    in the generator state; the slot to read is known a step ahead and is prefetched.
//   * post CTAs    : 128 CTAs, each owning 4 columns of the skip / postprocess1 matrices and 2 of
//                    postprocess2 RESIDENT in shared memory; z (all layers), relu(skip sum) and relu(post1) travel
//                    as tagged words through L2, every stage is one block-wide dot product.
//   * sampler CTA  : float64 softmax (model.py:619-621), temperature scaling (generate.py:229-233), np.cumsum /
//                    searchsorted('right') draw (np.random.choice), feeds the drawn id back to the chain head.
// wn_gen_run selects this kernel for streams == 1, commit, C == 32; everything else runs generator.cu.
#include <cooperative_groups.h>

#include "common.cuh"
#include "gen_common.h"
#include "kernels.h"

namespace wn {

namespace {
constexpr int C = 32;
constexpr int LPC = 8;          // layers (= warps) per chain CTA
constexpr int THREADS = 256;
constexpr int CS = 4;           // skip / post1 columns per post CTA
typedef unsigned long long u64;

struct LatArgs {
  GenArgs g;
  u64* comm;
  uint32_t tag_base;             // (launch_seq << 20); word tag = tag_base + step + 1
  int NC, NP, cq;                // chain CTAs, post CTAs, post2 columns per post CTA
  int causal_in_smem;
  long long* timeline;           // debug (wn_debug_timeline): %globaltimer stamps of step TL_STEP
};
constexpr int TL_STEP = 200;
__device__ __forceinline__ void tl_stamp(const LatArgs& a, int step, int slot) {
  if (a.timeline && step == TL_STEP) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
    a.timeline[slot] = (long long)g;
  }
}

// comm layout (u64 words)
__host__ __device__ inline int off_xg() { return 0; }                                   // [NC+1][32]
__host__ __device__ inline int off_zt(int NC) { return (NC + 1) * 32; }                 // [L*32]
__host__ __device__ inline int off_v0(int NC, int L) { return off_zt(NC) + L * 32; }    // [S]
__host__ __device__ inline int off_v1(int NC, int L, int S) { return off_v0(NC, L) + S; }
__host__ __device__ inline int off_lg(int NC, int L, int S) { return off_v1(NC, L, S) + S; }   // [Q]
__host__ __device__ inline int off_misc(int NC, int L, int S, int Q) { return off_lg(NC, L, S) + Q; }   // id, zdone, chain_done

__device__ __forceinline__ u64 pack(float v, uint32_t tag) { return ((u64)tag << 32) | (u64)__float_as_uint(v); }
__device__ __forceinline__ void st_gpu(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_gpu(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_gpu_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_gpu_f32(float* p, float v) { asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
// spin until the word carries `tag` (bounded: a protocol bug must trap, not hang the GPU)
__device__ __forceinline__ float wait_gpu(const u64* p, uint32_t tag) {
  for (uint32_t i = 0; i < (1u << 26); ++i) {
    const u64 v = ld_gpu(p);
    if ((uint32_t)(v >> 32) == tag) return __uint_as_float((uint32_t)v);
  }
  __trap();
  return 0.f;
}
__device__ __forceinline__ float wait_smem(const volatile u64* p, uint32_t tag) {
  for (uint32_t i = 0; i < (1u << 28); ++i) {
    const u64 v = *p;
    if ((uint32_t)(v >> 32) == tag) return __uint_as_float((uint32_t)v);
  }
  __trap();
  return 0.f;
}
// same, polling politely (long waits: one lane of a post / sampler CTA waiting for the chain)
__device__ __forceinline__ void wait_hint(const u64* p, uint32_t tag) {
  for (uint32_t i = 0; i < (1u << 24); ++i) {
    if ((uint32_t)(ld_gpu(p) >> 32) == tag) return;
    __nanosleep(32);
  }
  __trap();
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ------------------------------------------------------------------------------------------------
// chain: one warp = one layer
// ------------------------------------------------------------------------------------------------
__device__ void chain_warp(const LatArgs& a, int c, int w, volatile u64* xbuf /*[LPC+1][32]*/, float* priv /*[64]*/,
                           const float* causal_s) {
  const GenArgs& g = a.g;
  const int lane = threadIdx.x & 31;
  const int l = c * LPC + w;
  if (l >= g.L) return;
  const int d = g.dil[l];
  u64* comm = a.comm;
  u64* xg_in = comm + off_xg() + c * 32;
  u64* xg_out = comm + off_xg() + (c + 1) * 32;
  u64* zt = comm + off_zt(a.NC) + l * 32;
  u64* misc = comm + off_misc(a.NC, g.L, g.S, g.Q);
  float* ring = g.rings + (size_t)g.ring_off[l] * C;
  float* xs = priv;         // [32] layer input, broadcast source
  float* zs = priv + 32;    // [32] gated output, broadcast source

  // ---- this layer's weights -> registers (k-pairs for the packed FMA) ----
  float2 wfp[16], wfc[16], wgp[16], wgc[16], wd[16];
  {
    const float* F = g.filter + (size_t)l * 2 * C * C;   // [tap][k][n]
    const float* G = g.gate + (size_t)l * 2 * C * C;
    const float* D = g.dense + (size_t)l * C * C;        // [d][r]
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      wfp[kk] = make_float2(F[(2 * kk) * C + lane], F[(2 * kk + 1) * C + lane]);
      wfc[kk] = make_float2(F[(C + 2 * kk) * C + lane], F[(C + 2 * kk + 1) * C + lane]);
      wgp[kk] = make_float2(G[(2 * kk) * C + lane], G[(2 * kk + 1) * C + lane]);
      wgc[kk] = make_float2(G[(C + 2 * kk) * C + lane], G[(C + 2 * kk + 1) * C + lane]);
      wd[kk] = make_float2(D[(2 * kk) * C + lane], D[(2 * kk + 1) * C + lane]);
    }
  }
  float pbf = 0.f, pbg = 0.f, bd = 0.f;
  if (g.use_biases) {
    pbf = g.filter_bias[l * C + lane];
    pbg = g.gate_bias[l * C + lane];
    bd = g.dense_bias[l * C + lane];
  }
  if (g.G > 0 && g.gc_ids && g.gc_ids[0] >= 0 && g.gc_ids[0] < g.gc_card) {   // global conditioning: h . Wgc  (model.py:357-371)
    const float* e = g.gc_embedding + (size_t)g.gc_ids[0] * g.G;
    const float* wf = g.gc_filter + (size_t)l * g.G * C;
    const float* wg = g.gc_gate + (size_t)l * g.G * C;
    for (int k = 0; k < g.G; ++k) {
      pbf = fmaf(e[k], wf[k * C + lane], pbf);
      pbg = fmaf(e[k], wg[k * C + lane], pbg);
    }
  }
  const int step0 = g.hdr[1];
  int prev_id = g.hdr[0];
  const bool head = (l == 0);
  const bool sampling = g.uniforms != nullptr;

  for (int step = 0; step < g.n_steps; ++step) {
    const uint32_t T = a.tag_base + (uint32_t)step + 1u;
    const int slot = (step0 + step) % d;
    // ---- delay-line output x[t-d]: every lane reads the whole 128-byte slot; its half of the gated product
    //      is finished before this layer's input arrives ----
    float2 fa = make_float2(0.f, 0.f), ga = make_float2(0.f, 0.f);
    {
      const float* rp = ring + (size_t)slot * C;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 pv;
        asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(pv.x), "=f"(pv.y), "=f"(pv.z), "=f"(pv.w) : "l"(rp + 4 * j) : "memory");
        fa = ffma2(make_float2(pv.x, pv.y), wfp[2 * j], fa);
        ga = ffma2(make_float2(pv.x, pv.y), wgp[2 * j], ga);
        fa = ffma2(make_float2(pv.z, pv.w), wfp[2 * j + 1], fa);
        ga = ffma2(make_float2(pv.z, pv.w), wgp[2 * j + 1], ga);
      }
    }
    // ---- this layer's input ----
    float x_own;
    if (head) {
      int cur;
      if (g.forced) {
        cur = g.forced[step];
        if (step > 0) wait_gpu(misc + 2, T - 1);                         // previous step has left the chain
      } else if (step == 0) {
        cur = g.inputs[0];
      } else if (sampling) {
        cur = __float_as_int(wait_gpu(misc + 0, T - 1));                 // id drawn by the sampler for step-1
      } else {
        cur = g.inputs[0];
        wait_gpu(misc + 2, T - 1);
      }
      // causal layer: x = Wc[0][prev] + Wc[1][cur]   (model.py:341-346; zero history at the first step)
      const float* wc = a.causal_in_smem ? causal_s : g.causal;
      float v = 0.f;
      if (prev_id >= 0 && prev_id < g.Q) v += wc[(size_t)prev_id * C + lane];
      if (cur >= 0 && cur < g.Q) v += wc[(size_t)(g.Q + cur) * C + lane];
      x_own = v;
      prev_id = cur;
      if (lane == 0) { tl_stamp(a, step, 0); if (step == TL_STEP + 1 && a.timeline) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); a.timeline[15] = (long long)gt; } }
    } else if (w == 0) {
      x_own = wait_gpu(xg_in + lane, T);
    } else {
      x_own = wait_smem(xbuf + w * 32 + lane, T);
    }
    xs[lane] = x_own;
    __syncwarp();
    st_gpu_f32(ring + (size_t)slot * C + lane, x_own);      // push_ops: enqueue the layer input (model.py:461,482)
    float2 fb = make_float2(0.f, 0.f), gb = make_float2(0.f, 0.f), fc = fb, gc = gb;   // short dependent chains
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + 4 * j);
      fb = ffma2(make_float2(xv.x, xv.y), wfc[2 * j], fb);
      gb = ffma2(make_float2(xv.x, xv.y), wgc[2 * j], gb);
      fc = ffma2(make_float2(xv.z, xv.w), wfc[2 * j + 1], fc);
      gc = ffma2(make_float2(xv.z, xv.w), wgc[2 * j + 1], gc);
    }
    const float f = ((fa.x + fa.y) + pbf) + ((fb.x + fb.y) + (fc.x + fc.y));
    const float gg = ((ga.x + ga.y) + pbg) + ((gb.x + gb.y) + (gc.x + gc.y));
    const float z = tanh_fast(f) * sigmoid_fast(gg);
    zs[lane] = z;
    __syncwarp();
    if (l + 1 < g.L) {      // the last layer's dense output is discarded by the reference (model.py:377-380)
      float2 oa = make_float2(0.f, 0.f), ob = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 zv = *reinterpret_cast<const float4*>(zs + 4 * j);
        oa = ffma2(make_float2(zv.x, zv.y), wd[2 * j], oa);
        ob = ffma2(make_float2(zv.z, zv.w), wd[2 * j + 1], ob);
      }
      const float x_out = x_own + bd + ((oa.x + oa.y) + (ob.x + ob.y));
      if (w + 1 < LPC) xbuf[(w + 1) * 32 + lane] = pack(x_out, T);
      else { st_gpu(xg_out + lane, pack(x_out, T)); if (lane == 0) tl_stamp(a, step, 1 + c); }
    }
    st_gpu(zt + lane, pack(z, T));
    if (l == g.L - 1) {
      __syncwarp();
      if (lane == 31) {
        st_gpu(misc + 1, pack(0.f, T));     // hint for the post CTAs: the last z of this step is on its way
        st_gpu(misc + 2, pack(0.f, T));     // the chain is free for the next forced step
        tl_stamp(a, step, 1 + c);
      }
    }
    __syncwarp();      // xs / zs are rewritten next step
  }
  if (head && lane == 0) {   // generator header: model.py push of the causal queue + step counter
    g.hdr[0] = prev_id;
    g.hdr[1] = step0 + g.n_steps;
    g.hdr[3] = 0;
  }
}

// Loads words p[tid], p[tid + 256], ... (n <= 8 * 256) in one batch and spins on the batch until every word
// carries `tag`: one L2 round trip when the producers are done, instead of one per word.
template <int MAXW>
__device__ __forceinline__ void wait_batch(const u64* p, int n, uint32_t tag, float (&v)[MAXW]) {
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    u64 w[MAXW];
#pragma unroll
    for (int j = 0; j < MAXW; ++j) {
      const int k = threadIdx.x + j * THREADS;
      w[j] = (k < n) ? ld_gpu(p + k) : ((u64)tag << 32);
    }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < MAXW; ++j) {
      ok = ok && ((uint32_t)(w[j] >> 32) == tag);
      v[j] = __uint_as_float((uint32_t)w[j]);
    }
    if (ok) return;
  }
  __trap();
}

// block-wide sum of a float4 per thread (256 threads); result valid in every thread
__device__ __forceinline__ float4 block_sum4(float4 v, float4* red /*[2][8]*/, int& flip) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  float4* r = red + 8 * flip;      // alternating scratch: the barrier of call i+1 protects the buffer of call i
  flip ^= 1;
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  float4 t = r[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) {
    t.x += r[w].x; t.y += r[w].y; t.z += r[w].z; t.w += r[w].w;
  }
  return t;
}

// ------------------------------------------------------------------------------------------------
// post-processing CTA p: columns [4p, 4p+4) of skip / postprocess1, [cq*p, cq*p+cq) of postprocess2
// ------------------------------------------------------------------------------------------------
__device__ void post_cta(const LatArgs& a, int p, float* sm) {
  const GenArgs& g = a.g;
  const int tid = threadIdx.x;
  const int LD = g.L * C, S = g.S, Q = g.Q, cq = a.cq;
  float4* ws = reinterpret_cast<float4*>(sm);            // [LD]  skip weights, 4 columns
  float4* w1 = ws + LD;                                   // [S]
  float4* w2 = w1 + S;                                    // [S]   (cq <= 4 columns used)
  float4* red = w2 + S;                                   // [8]
  __shared__ float bias_s[12];                            // skip-bias sum | post1 bias | post2 bias of my columns
  for (int k = tid; k < LD; k += THREADS) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = g.skip + (size_t)k * S + CS * p;
    if (CS * p + 0 < S) v.x = src[0];
    if (CS * p + 1 < S) v.y = src[1];
    if (CS * p + 2 < S) v.z = src[2];
    if (CS * p + 3 < S) v.w = src[3];
    ws[k] = v;
  }
  for (int k = tid; k < S; k += THREADS) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = g.post1 + (size_t)k * S + CS * p;
    if (CS * p + 0 < S) v.x = src[0];
    if (CS * p + 1 < S) v.y = src[1];
    if (CS * p + 2 < S) v.z = src[2];
    if (CS * p + 3 < S) v.w = src[3];
    w1[k] = v;
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* s2 = g.post2 + (size_t)k * Q + cq * p;
    if (0 < cq && cq * p + 0 < Q) u.x = s2[0];
    if (1 < cq && cq * p + 1 < Q) u.y = s2[1];
    if (2 < cq && cq * p + 2 < Q) u.z = s2[2];
    if (3 < cq && cq * p + 3 < Q) u.w = s2[3];
    w2[k] = u;
  }
  if (tid < 12) {
    float v = 0.f;
    if (g.use_biases) {
      const int j = tid & 3;
      if (tid < 4) {
        if (CS * p + j < S)
          for (int l = 0; l < g.L; ++l) v += g.skip_bias[(size_t)l * S + CS * p + j];   // model.py:430 sum of skips
      } else if (tid < 8) {
        if (CS * p + j < S) v = g.post1_bias[CS * p + j];
      } else {
        if (j < cq && cq * p + j < Q) v = g.post2_bias[cq * p + j];
      }
    }
    bias_s[tid] = v;
  }
  __syncthreads();
  u64* comm = a.comm;
  const u64* zt = comm + off_zt(a.NC);
  u64* v0t = comm + off_v0(a.NC, g.L);
  u64* v1t = comm + off_v1(a.NC, g.L, S);
  u64* lgt = comm + off_lg(a.NC, g.L, S);
  const u64* misc = comm + off_misc(a.NC, g.L, S, Q);
  const bool sampling = g.uniforms != nullptr;
  int flip = 0;

  for (int step = 0; step < g.n_steps; ++step) {
    if (!sampling && step != g.n_steps - 1) continue;       // priming: only the last distribution is needed
    const uint32_t T = a.tag_base + (uint32_t)step + 1u;
    if (tid == 0) wait_hint(misc + 1, T);
    __syncthreads();
    if (p == 0 && tid == 0) tl_stamp(a, step, 8);
    // ---- skip sum -> relu            (model.py:505-507)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      float zv[8];
      wait_batch<8>(zt, LD, T, zv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = tid + j * THREADS;
        if (k < LD) {
          const float4 w = ws[k];
          acc.x = fmaf(zv[j], w.x, acc.x); acc.y = fmaf(zv[j], w.y, acc.y); acc.z = fmaf(zv[j], w.z, acc.z); acc.w = fmaf(zv[j], w.w, acc.w);
        }
      }
    }
    acc = block_sum4(acc, red, flip);
    if (tid < CS && CS * p + tid < S) {
      const float v = (tid == 0 ? acc.x : tid == 1 ? acc.y : tid == 2 ? acc.z : acc.w) + bias_s[tid];
      st_gpu(v0t + CS * p + tid, pack(fmaxf(v, 0.f), T));
      if (p == 0 && tid == 0) tl_stamp(a, step, 9);
    }
    // ---- postprocess1 -> relu        (model.py:508-511)
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      float xv[4];
      wait_batch<4>(v0t, S, T, xv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = tid + j * THREADS;
        if (k < S) {
          const float4 w = w1[k];
          acc.x = fmaf(xv[j], w.x, acc.x); acc.y = fmaf(xv[j], w.y, acc.y); acc.z = fmaf(xv[j], w.z, acc.z); acc.w = fmaf(xv[j], w.w, acc.w);
        }
      }
    }
    acc = block_sum4(acc, red, flip);
    if (tid < CS && CS * p + tid < S) {
      const float v = (tid == 0 ? acc.x : tid == 1 ? acc.y : tid == 2 ? acc.z : acc.w) + bias_s[4 + tid];
      st_gpu(v1t + CS * p + tid, pack(fmaxf(v, 0.f), T));
      if (p == 0 && tid == 0) tl_stamp(a, step, 10);
    }
    // ---- postprocess2                (model.py:512-514)
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      float xv[4];
      wait_batch<4>(v1t, S, T, xv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = tid + j * THREADS;
        if (k < S) {
          const float4 w = w2[k];
          acc.x = fmaf(xv[j], w.x, acc.x); acc.y = fmaf(xv[j], w.y, acc.y); acc.z = fmaf(xv[j], w.z, acc.z); acc.w = fmaf(xv[j], w.w, acc.w);
        }
      }
    }
    acc = block_sum4(acc, red, flip);
    if (tid < cq && cq * p + tid < Q) {
      const float v = (tid == 0 ? acc.x : tid == 1 ? acc.y : tid == 2 ? acc.z : acc.w) + bias_s[8 + tid];
      st_gpu(lgt + cq * p + tid, pack(v, T));
      if (p == 0 && tid == 0) tl_stamp(a, step, 11);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// sampler CTA: float64 softmax, temperature, inverse-cdf draw
// ------------------------------------------------------------------------------------------------
// Handles streams s_first, s_first + s_stride, ... < NS (single-stream kernel: 0, 1, 1); comm words per stream: wps.
__device__ void sampler_cta(const LatArgs& a, float* sm, int s_first = 0, int s_stride = 1, int NS = 1, int wps = 0) {
  const GenArgs& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Q = g.Q;
  double* dsm = reinterpret_cast<double*>(sm);     // [Q] exp values
  float* ps = reinterpret_cast<float*>(dsm + Q);   // [Q] float32 probabilities
  __shared__ double redd[2][8];
  __shared__ float redf[2][8];
  const bool sampling = g.uniforms != nullptr;
  const int per = (Q + 31) / 32;                   // cdf elements per lane of warp 0 (Q <= 1024)
  int flip = 0;

  for (int step = 0; step < g.n_steps; ++step)
  for (int s = s_first; s < NS; s += s_stride) {
    if (!sampling && step != g.n_steps - 1) continue;
    u64* comm = a.comm + (size_t)s * wps;
    const u64* lgt = comm + off_lg(a.NC, g.L, g.S);
    u64* misc = comm + off_misc(a.NC, g.L, g.S, Q);
    const uint32_t T = a.tag_base + (uint32_t)step + 1u;
    const double u = sampling ? g.uniforms[(size_t)s * g.n_steps + step] : 0.0;     // fetched while the chain is still running
    if (tid == 0) wait_hint(misc + 1, T);
    __syncthreads();
    // float64 softmax of the logits, cast back to float32   (model.py:619-621)
    float lg[4];
    wait_batch<4>(lgt, Q, T, lg);
    float mxf = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (tid + j * THREADS < Q) mxf = fmaxf(mxf, lg[j]);
    mxf = warp_max(mxf);
    if (lane == 0) redf[flip][warp] = mxf;
    __syncthreads();
    mxf = redf[flip][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mxf = fmaxf(mxf, redf[flip][w]);
    const double mx = (double)mxf;
    double e[4];
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      e[j] = (tid + j * THREADS < Q) ? exp((double)lg[j] - mx) : 0.0;
      sum += e[j];
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) redd[flip][warp] = sum;
    __syncthreads();
    sum = redd[flip][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) sum += redd[flip][w];
    flip ^= 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + j * THREADS;
      if (i < Q) ps[i] = (float)(e[j] / sum);
    }
    if (g.temperature != 1.0f) {     // generate.py:229-233, float32 log-space
      __syncthreads();
      float m2 = -INFINITY;
      for (int i = tid; i < Q; i += THREADS) {
        const float q = logf(ps[i]) / g.temperature;
        ps[i] = q;
        m2 = fmaxf(m2, q);
      }
      m2 = warp_max(m2);
      if (lane == 0) redf[flip][warp] = m2;
      __syncthreads();
      m2 = redf[flip][0];
      for (int w = 1; w < 8; ++w) m2 = fmaxf(m2, redf[flip][w]);
      float s2 = 0.f;
      for (int i = tid; i < Q; i += THREADS) s2 += expf(ps[i] - m2);
      s2 = warp_sum(s2);
      flip ^= 1;
      if (lane == 0) redf[flip][warp] = s2;
      __syncthreads();
      s2 = redf[flip][0];
      for (int w = 1; w < 8; ++w) s2 += redf[flip][w];
      const float lse = m2 + logf(s2);
      for (int i = tid; i < Q; i += THREADS) ps[i] = expf(ps[i] - lse);
    }
    __syncthreads();      // publishes ps[] to warp 0
    if (tid == 0) tl_stamp(a, step, 12);
    if (g.proba_out && step == g.n_steps - 1)
      for (int i = tid; i < Q; i += THREADS) g.proba_out[(size_t)s * Q + i] = ps[i];
    if (sampling && warp == 0) {
      // np.random.choice: cdf = cumsum(p) (sequential float64 adds), normalise by the last, searchsorted right.
      // A warp scan computes the same sums in another association: every partial sum differs from np.cumsum's
      // by < Q * 2^-53 relative (all terms are positive), so  cdf[i] / total <= u  has the same truth value unless
      // cdf[i] lies within 1e-12 (relative) of u * total.  Only in that case (p ~ 1e-10 per sample) lane 0
      // repeats the sum sequentially and the exact quotients are compared.
      int cnt = 0;
      double t = 0.0;
      for (int j = 0; j < per; ++j) {
        const int i = lane * per + j;
        t += (i < Q) ? (double)ps[i] : 0.0;
      }
      double incl = t;
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const double total = __shfl_sync(0xffffffffu, incl, 31);
      const double thr = u * total, eps = thr * 1e-12;
      double run = incl - t;
      bool ambiguous = false;
      for (int j = 0; j < per; ++j) {
        const int i = lane * per + j;
        if (i < Q) {
          run += (double)ps[i];
          const double dlt = run - thr;
          ambiguous = ambiguous || (fabs(dlt) <= eps);
          cnt += (dlt < 0.0) ? 1 : 0;
        }
      }
      if (__any_sync(0xffffffffu, ambiguous)) {
        if (lane == 0) {
          double r2 = 0.0;
          for (int i = 0; i < Q; ++i) { r2 += (double)ps[i]; dsm[i] = r2; }
        }
        __syncwarp();
        const double tot2 = dsm[Q - 1];
        cnt = 0;
        for (int i = lane; i < Q; i += 32) cnt += (dsm[i] / tot2 <= u) ? 1 : 0;
      }
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0) {
        const int drawn = cnt < Q ? cnt : Q - 1;
        st_gpu(misc + 0, pack(__int_as_float(drawn), T));      // feeds the chain head of the next step
        g.samples_out[(size_t)s * g.n_steps + step] = drawn;
        tl_stamp(a, step, 13);
      }
    }
    __syncthreads();      // ps / dsm are rewritten by the next (stream, step)
  }
}

__global__ void __launch_bounds__(THREADS, 1) generator_lat_kernel(LatArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x;
  if (b < a.NC) {
    volatile u64* xbuf = reinterpret_cast<volatile u64*>(sm);                 // [LPC+1][32] tagged words
    float* priv = sm + 2 * (LPC + 1) * 32 + (threadIdx.x >> 5) * 64;          // per warp [64]
    float* causal_s = sm + 2 * (LPC + 1) * 32 + LPC * 64;
    for (int i = threadIdx.x; i < (LPC + 1) * 32; i += THREADS) xbuf[i] = 0ull;
    if (b == 0 && a.causal_in_smem)
      for (int i = threadIdx.x; i < 2 * a.g.Q * C; i += THREADS) causal_s[i] = a.g.causal[i];
    __syncthreads();
    chain_warp(a, b, threadIdx.x >> 5, xbuf, priv, causal_s);
  } else if (b == a.NC) {
    sampler_cta(a, sm);
  } else if (b < a.NC + 1 + a.NP) {
    post_cta(a, b - a.NC - 1, sm);
  }
}

// =================================================================================================
// Pipelined form: NS independent streams ride the SAME layer-per-warp chain.  A layer warp (weights in registers) takes
// the streams round robin -- (s = 0, t), (s = 1, t), ..., (s = NS-1, t), (s = 0, t+1) -- so while stream s waits for its
// sample to come back from the post-processing CTAs and its sampler, the warp works on the other streams: one hop
// (~260 ns) per stream and layer, i.e. up to ~19 us / 0.26 us ~ 70 streams at the single-stream latency.  Every tagged
// word, delay line, uniform and output exists once per stream (comm + s * words_per_stream); the post CTAs run their
// three dot products phase by phase over all streams, and the streams are dealt out to NSAMP sampler CTAs.
// =================================================================================================
constexpr int NS_MAX = 32;       // streams per launch (shared memory: per-stream hand-off slots and conditioning rows)
constexpr int NSAMP = 8;         // sampler CTAs

struct PipeArgs {
  LatArgs a;
  int NS, wps;                   // streams, comm words per stream
  int pb_in_smem;                // per-stream conditioning rows (global conditioning) in shared memory
};

__device__ void chain_warp_ms(const PipeArgs& pa, int c, int w, volatile u64* xbuf /*[NS][LPC+1][32]*/, float* priv /*[64 + 4*32]*/,
                              const float* causal_s, float* pb_s /*[LPC][NS][64] or null*/, int* st_s /*[NS] step0, [NS] prev id*/) {
  const LatArgs& a = pa.a;
  const GenArgs& g = a.g;
  const int lane = threadIdx.x & 31;
  const int l = c * LPC + w;
  const int NS = pa.NS;
  if (l >= g.L) return;
  const int d = g.dil[l];
  float* xs = priv;         // [32] layer input, broadcast source
  float* zs = priv + 32;    // [32] gated output, broadcast source
  float* rb = priv + 64;    // [4][32] delay-line slots of the next items (cp.async ring)

  float2 wfp[16], wfc[16], wgp[16], wgc[16], wd[16];
  {
    const float* F = g.filter + (size_t)l * 2 * C * C;   // [tap][k][n]
    const float* G = g.gate + (size_t)l * 2 * C * C;
    const float* D = g.dense + (size_t)l * C * C;        // [d][r]
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      wfp[kk] = make_float2(F[(2 * kk) * C + lane], F[(2 * kk + 1) * C + lane]);
      wfc[kk] = make_float2(F[(C + 2 * kk) * C + lane], F[(C + 2 * kk + 1) * C + lane]);
      wgp[kk] = make_float2(G[(2 * kk) * C + lane], G[(2 * kk + 1) * C + lane]);
      wgc[kk] = make_float2(G[(C + 2 * kk) * C + lane], G[(C + 2 * kk + 1) * C + lane]);
      wd[kk] = make_float2(D[(2 * kk) * C + lane], D[(2 * kk + 1) * C + lane]);
    }
  }
  float pbf0 = 0.f, pbg0 = 0.f, bd = 0.f;
  if (g.use_biases) {
    pbf0 = g.filter_bias[l * C + lane];
    pbg0 = g.gate_bias[l * C + lane];
    bd = g.dense_bias[l * C + lane];
  }
  float* pbw = pb_s ? pb_s + (size_t)w * NS * 64 : nullptr;
  if (pbw) {      // global conditioning: h_s . Wgc per stream  (model.py:357-371)
    const float* wf = g.gc_filter + (size_t)l * g.G * C;
    const float* wg = g.gc_gate + (size_t)l * g.G * C;
    for (int s = 0; s < NS; ++s) {
      float vf = pbf0, vg = pbg0;
      const int id = g.gc_ids[s];
      if (id >= 0 && id < g.gc_card) {
        const float* e = g.gc_embedding + (size_t)id * g.G;
        for (int k = 0; k < g.G; ++k) {
          vf = fmaf(e[k], wf[k * C + lane], vf);
          vg = fmaf(e[k], wg[k * C + lane], vg);
        }
      }
      pbw[s * 64 + lane] = vf;
      pbw[s * 64 + 32 + lane] = vg;
    }
    __syncwarp();
  }
  const bool head = (l == 0);
  const bool sampling = g.uniforms != nullptr;
  const size_t ring_stride = (size_t)g.sum_d * C;
  const float* ring0 = g.rings + (size_t)g.ring_off[l] * C;
  // ring slot of (stream s, step) = (step0_s + step) mod d, without a division per item: smod[s] = step0_s mod d once,
  // step mod d carried along
  float* smod_f = priv + 192;      // [NS] ints, per warp
  int* smod = reinterpret_cast<int*>(smod_f);
  for (int s = lane; s < NS; s += 32) smod[s] = st_s[s] % d;
  __syncwarp();
  auto slot_of = [&](int s, int step_mod) { const int v = smod[s] + step_mod; return v >= d ? v - d : v; };
  // The delay-line slot of an item is requested PD items ahead (an L2 round trip is ~1 us, an item ~0.3 us).  PD < NS: with
  // d = 1 the slot of (s, t) is the one (s, t-1) enqueued NS items earlier, and the request must follow that store.
  const int PD = NS > 3 ? 3 : NS - 1;
  int p_s = 0, p_step = 0, p_mod = 0, p_buf = 0;      // the next item to request
  auto prefetch_next = [&]() {      // lanes 0-7 move 16 B each into rb[p_buf]
    if (p_step < g.n_steps) {
      const int slot = slot_of(p_s, p_mod);
      if (lane < 8) cp_async16(rb + p_buf * 32 + 4 * lane, ring0 + (size_t)p_s * ring_stride + (size_t)slot * C + 4 * lane, true);
    }
    cp_async_commit();
    p_buf = (p_buf + 1) & 3;
    if (++p_s == NS) { p_s = 0; ++p_step; if (++p_mod == d) p_mod = 0; }
  };
  for (int k = 0; k < PD; ++k) prefetch_next();
  // first warp of a chain CTA (input through L2) and the head (sampled id through L2): the words of the next FW items are
  // requested while this item is computed (a ring of FW registers): a consumer that looks for item k+1 only when it is done
  // with item k pays an L2 round trip per item, and the head then caps the whole chain at one item per round trip
  constexpr int FW = 4;
  const bool l2_in = (w == 0);
  u64 in_ring[FW];
#pragma unroll
  for (int j = 0; j < FW; ++j) in_ring[j] = 0ull;
  auto in_word = [&](int s) -> const u64* {
    u64* comm = a.comm + (size_t)s * pa.wps;
    if (!head) return comm + off_xg() + c * 32 + lane;
    return comm + off_misc(a.NC, g.L, g.S, g.Q) + ((sampling && !g.forced) ? 0 : 2);
  };
  int it = 0, step_mod = 0;
  for (int step = 0; step < g.n_steps; ++step, step_mod = (step_mod + 1 == d ? 0 : step_mod + 1)) {
    const uint32_t T = a.tag_base + (uint32_t)step + 1u;
    for (int s = 0; s < NS; ++s, ++it) {
      u64* comm = a.comm + (size_t)s * pa.wps;
      u64* misc = comm + off_misc(a.NC, g.L, g.S, g.Q);
      const int buf = it & 3;
      prefetch_next();
      if (PD == 3) cp_async_wait<3>();
      else if (PD == 2) cp_async_wait<2>();
      else cp_async_wait<1>();
      __syncwarp();
      const int slot = slot_of(s, step_mod);
      float* ring = const_cast<float*>(ring0) + (size_t)s * ring_stride;
      // ---- delay-line output x[t-d] times the past-tap weights (done before this layer's input is looked at) ----
      float2 fa = make_float2(0.f, 0.f), ga = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 pv = *reinterpret_cast<const float4*>(rb + buf * 32 + 4 * j);
        fa = ffma2(make_float2(pv.x, pv.y), wfp[2 * j], fa);
        ga = ffma2(make_float2(pv.x, pv.y), wgp[2 * j], ga);
        fa = ffma2(make_float2(pv.z, pv.w), wfp[2 * j + 1], fa);
        ga = ffma2(make_float2(pv.z, pv.w), wgp[2 * j + 1], ga);
      }
      // ---- this layer's input ----
      // (the L2 word of this item was requested up to FW items ago; the tag it must carry: T, or T-1 for the head's feedback
      // words; a stale word just means polling as before)
      const u64 in_cur = in_ring[0];
      if (l2_in) {
#pragma unroll
        for (int j = 0; j + 1 < FW; ++j) in_ring[j] = in_ring[j + 1];
        // the word FW items ahead; with fewer than FW + 1 streams it would be this stream's own NEXT step, whose word is
        // still this step's: then the request is useless but harmless (tag mismatch -> polled when needed)
        int fs = s + FW;
        while (fs >= NS) fs -= NS;
        in_ring[FW - 1] = ld_gpu(in_word(fs));
      }
      float x_own;
      if (head) {
        int cur;
        if (g.forced) {
          cur = g.forced[(size_t)s * g.n_steps + step];
          if (step > 0 && (uint32_t)(in_cur >> 32) != T - 1) wait_gpu(misc + 2, T - 1);      // previous step of this stream has left the chain
        } else if (step == 0) {
          cur = g.inputs[s];
        } else if (sampling) {                                             // id drawn by the sampler for step-1
          cur = ((uint32_t)(in_cur >> 32) == T - 1) ? (int)(uint32_t)in_cur : __float_as_int(wait_gpu(misc + 0, T - 1));
        } else {
          cur = g.inputs[s];
          if ((uint32_t)(in_cur >> 32) != T - 1) wait_gpu(misc + 2, T - 1);
        }
        // causal layer: x = Wc[0][prev] + Wc[1][cur]   (model.py:341-346; zero history at the first step)
        const float* wc = a.causal_in_smem ? causal_s : g.causal;
        const int prev_id = st_s[NS + s];
        float v = 0.f;
        if (prev_id >= 0 && prev_id < g.Q) v += wc[(size_t)prev_id * C + lane];
        if (cur >= 0 && cur < g.Q) v += wc[(size_t)(g.Q + cur) * C + lane];
        x_own = v;
        __syncwarp();
        if (lane == 0) st_s[NS + s] = cur;
      } else if (w == 0) {
        x_own = ((uint32_t)(in_cur >> 32) == T) ? __uint_as_float((uint32_t)in_cur) : wait_gpu(comm + off_xg() + c * 32 + lane, T);
      } else {
        x_own = wait_smem(xbuf + ((size_t)s * (LPC + 1) + w) * 32 + lane, T);
      }
      xs[lane] = x_own;
      __syncwarp();
      st_gpu_f32(ring + (size_t)slot * C + lane, x_own);      // push_ops: enqueue the layer input (model.py:461,482)
      float2 fb = make_float2(0.f, 0.f), gb = make_float2(0.f, 0.f), fc = fb, gc = gb;   // short dependent chains
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + 4 * j);
        fb = ffma2(make_float2(xv.x, xv.y), wfc[2 * j], fb);
        gb = ffma2(make_float2(xv.x, xv.y), wgc[2 * j], gb);
        fc = ffma2(make_float2(xv.z, xv.w), wfc[2 * j + 1], fc);
        gc = ffma2(make_float2(xv.z, xv.w), wgc[2 * j + 1], gc);
      }
      const float pbf = pbw ? pbw[s * 64 + lane] : pbf0, pbg = pbw ? pbw[s * 64 + 32 + lane] : pbg0;
      const float f = ((fa.x + fa.y) + pbf) + ((fb.x + fb.y) + (fc.x + fc.y));
      const float gg = ((ga.x + ga.y) + pbg) + ((gb.x + gb.y) + (gc.x + gc.y));
      const float z = tanh_fast(f) * sigmoid_fast(gg);
      zs[lane] = z;
      __syncwarp();
      if (l + 1 < g.L) {      // the last layer's dense output is discarded by the reference (model.py:377-380)
        float2 oa = make_float2(0.f, 0.f), ob = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 zv = *reinterpret_cast<const float4*>(zs + 4 * j);
          oa = ffma2(make_float2(zv.x, zv.y), wd[2 * j], oa);
          ob = ffma2(make_float2(zv.z, zv.w), wd[2 * j + 1], ob);
        }
        const float x_out = x_own + bd + ((oa.x + oa.y) + (ob.x + ob.y));
        if (w + 1 < LPC) xbuf[((size_t)s * (LPC + 1) + w + 1) * 32 + lane] = pack(x_out, T);
        else st_gpu(comm + off_xg() + (c + 1) * 32 + lane, pack(x_out, T));
      }
      st_gpu(comm + off_zt(a.NC) + l * 32 + lane, pack(z, T));
      if (l == g.L - 1) {
        __syncwarp();
        if (lane == 31) {
          st_gpu(misc + 1, pack(0.f, T));     // hint for the post CTAs: the last z of this step is on its way
          st_gpu(misc + 2, pack(0.f, T));     // the chain is free for the next forced step of this stream
        }
      }
      __syncwarp();      // xs / zs are rewritten by the next item
    }
  }
  cp_async_wait<0>();
  if (head && lane == 0) {   // generator headers: model.py push of the causal queue + step counter
    for (int s = 0; s < NS; ++s) {
      g.hdr[s * 4 + 0] = st_s[NS + s];
      g.hdr[s * 4 + 1] = st_s[s] + g.n_steps;
      g.hdr[s * 4 + 3] = 0;
    }
  }
}

// Post-processing CTAs of the pipelined form.  With every CTA walking through ALL streams (4 columns each, like the
// single-stream kernel) 32 streams cost 24 phase passes of ~2.3 us per step -- the bottleneck of the whole kernel (the
// chain needs ~0.5 us per stream and step).  Here the NP = 128 CTAs form NG = 4 groups of 32; a group serves every fourth
// batch of SB = 4 streams, and a CTA owns 16 columns of skip / postprocess1 (8 of postprocess2): six passes per step.
// One pass (one phase, one batch): the batch's tagged input words are polled together and staged in shared memory; thread
// (column c, k-slice j) multiplies its slice of the inputs of all SB streams with its weight column; the 16 k-slices are
// summed with four shuffles.
constexpr int SB = 4;            // streams per batch
constexpr int NG = 4;            // groups of post CTAs (batches are dealt out round robin)
constexpr int PCW = 16;          // columns of skip / postprocess1 per post CTA

template <int MAXW>
__device__ __forceinline__ void post_pass(const u64* in0, u64* out0, size_t wps, int nb, int n_in, const float* wsm /*[n_in][ncols]*/,
                                          int ncols, const float* bias, int col0, int n_out, bool relu, uint32_t T, float* stage /*[SB][n_in]*/) {
  const int tid = threadIdx.x, lane = tid & 31;
  {
    float xv[SB][MAXW];
    for (uint32_t it = 0;; ++it) {
      bool ok = true;
#pragma unroll
      for (int b = 0; b < SB; ++b) {
        if (b < nb) {
#pragma unroll
          for (int j = 0; j < MAXW; ++j) {
            const int k = tid + j * THREADS;
            const u64 w = (k < n_in) ? ld_gpu(in0 + (size_t)b * wps + k) : ((u64)T << 32);
            ok = ok && ((uint32_t)(w >> 32) == T);
            xv[b][j] = __uint_as_float((uint32_t)w);
          }
        }
      }
      if (ok) break;
      if (it > (1u << 24)) __trap();
    }
    __syncthreads();      // (the previous pass has finished reading the staging area)
#pragma unroll
    for (int b = 0; b < SB; ++b) {
#pragma unroll
      for (int j = 0; j < MAXW; ++j) {
        const int k = tid + j * THREADS;
        if (k < n_in) stage[b * n_in + k] = (b < nb) ? xv[b][j] : 0.f;
      }
    }
  }
  __syncthreads();
  // thread = (k-slice, column): lanes 0-15 and 16-31 of a warp hold two consecutive k-slices of the 16 columns
  const int c = tid & 15, ks = tid >> 4;      // 16 k-slices
  float acc[SB];
#pragma unroll
  for (int b = 0; b < SB; ++b) acc[b] = 0.f;
  if (c < ncols) {
#pragma unroll 4
    for (int k = ks; k < n_in; k += 16) {      // (unrolled: the shared-memory loads of four k in flight)
      const float w = wsm[k * ncols + c];
#pragma unroll
      for (int b = 0; b < SB; ++b) acc[b] = fmaf(stage[b * n_in + k], w, acc[b]);
    }
  }
  // sum over the 16 k-slices: lanes l and l ^ 16 (two slices of a warp), then the eight warps through shared memory
#pragma unroll
  for (int b = 0; b < SB; ++b) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], 16);
  __shared__ float part[8][SB][16];
  if (lane < 16) {
#pragma unroll
    for (int b = 0; b < SB; ++b) part[tid >> 5][b][lane] = acc[b];
  }
  __syncthreads();
  if (tid < SB * 16) {
    const int b = tid >> 4, cc = tid & 15;
    if (b < nb && cc < ncols && col0 + cc < n_out) {
      float v = bias[cc];
#pragma unroll
      for (int w = 0; w < 8; ++w) v += part[w][b][cc];
      if (relu) v = fmaxf(v, 0.f);
      st_gpu(out0 + (size_t)b * wps + col0 + cc, pack(v, T));
    }
  }
}

__device__ void post_cta_ms(const PipeArgs& pa, int p, float* sm) {
  const LatArgs& a = pa.a;
  const GenArgs& g = a.g;
  const int tid = threadIdx.x;
  const int LD = g.L * C, S = g.S, Q = g.Q;
  const int per_group = a.NP / NG;             // CTAs of a group
  const int grp = p / per_group, pc = p % per_group;
  const int ncq = (Q + per_group - 1) / per_group;      // postprocess2 columns per CTA (<= PCW)
  float* ws = sm;                               // [LD][PCW] skip weights of my columns
  float* w1 = ws + (size_t)LD * PCW;            // [S][PCW]
  float* w2 = w1 + (size_t)S * PCW;             // [S][ncq]
  float* stage = w2 + (size_t)S * ncq;          // [SB][max(LD, S)]
  __shared__ float bias_s[3][PCW];              // skip-bias sum | post1 bias | post2 bias of my columns
  for (int i = tid; i < LD * PCW; i += THREADS) {
    const int k = i / PCW, c = i % PCW, col = PCW * pc + c;
    ws[i] = col < S ? g.skip[(size_t)k * S + col] : 0.f;
  }
  for (int i = tid; i < S * PCW; i += THREADS) {
    const int k = i / PCW, c = i % PCW, col = PCW * pc + c;
    w1[i] = col < S ? g.post1[(size_t)k * S + col] : 0.f;
  }
  for (int i = tid; i < S * ncq; i += THREADS) {
    const int k = i / ncq, c = i % ncq, col = ncq * pc + c;
    w2[i] = col < Q ? g.post2[(size_t)k * Q + col] : 0.f;
  }
  if (tid < 3 * PCW) {
    const int which = tid / PCW, c = tid % PCW;
    float v = 0.f;
    if (g.use_biases) {
      if (which == 0) {
        if (PCW * pc + c < S)
          for (int l = 0; l < g.L; ++l) v += g.skip_bias[(size_t)l * S + PCW * pc + c];   // model.py:430 sum of skips
      } else if (which == 1) {
        if (PCW * pc + c < S) v = g.post1_bias[PCW * pc + c];
      } else {
        if (c < ncq && ncq * pc + c < Q) v = g.post2_bias[ncq * pc + c];
      }
    }
    bias_s[which][c] = v;
  }
  __syncthreads();
  const bool sampling = g.uniforms != nullptr;
  const size_t wps = pa.wps;
  const int nbatch = (pa.NS + SB - 1) / SB;
  for (int step = 0; step < g.n_steps; ++step) {
    if (!sampling && step != g.n_steps - 1) continue;       // priming: only the last distribution is needed
    const uint32_t T = a.tag_base + (uint32_t)step + 1u;
    // my batches: grp, grp + NG, ...  Software pipeline over them: iteration k runs the skip sum of my k-th batch,
    // postprocess1 of the (k-1)-th and postprocess2 of the (k-2)-th -- the L2 round trip between two phases of a batch (all
    // CTAs of the group publish their columns) is covered by the other batches.
    const int mine = grp < nbatch ? (nbatch - grp + NG - 1) / NG : 0;
    for (int k = 0; k < mine + 2; ++k) {
      if (k < mine) {                          // skip sum -> relu            (model.py:505-507)
        const int s0 = (grp + k * NG) * SB, nb = min(SB, pa.NS - s0);
        u64* comm = a.comm + (size_t)s0 * wps;
        if (tid == 0) wait_hint(comm + (size_t)(nb - 1) * wps + off_misc(a.NC, g.L, S, Q) + 1, T);      // (streams leave the chain in order)
        __syncthreads();
        post_pass<8>(comm + off_zt(a.NC), comm + off_v0(a.NC, g.L), wps, nb, LD, ws, PCW, bias_s[0], PCW * pc, S, true, T, stage);
      }
      if (k >= 1 && k - 1 < mine) {            // postprocess1 -> relu        (model.py:508-511)
        const int s0 = (grp + (k - 1) * NG) * SB, nb = min(SB, pa.NS - s0);
        u64* comm = a.comm + (size_t)s0 * wps;
        post_pass<4>(comm + off_v0(a.NC, g.L), comm + off_v1(a.NC, g.L, S), wps, nb, S, w1, PCW, bias_s[1], PCW * pc, S, true, T, stage);
      }
      if (k >= 2 && k - 2 < mine) {            // postprocess2                (model.py:512-514)
        const int s0 = (grp + (k - 2) * NG) * SB, nb = min(SB, pa.NS - s0);
        u64* comm = a.comm + (size_t)s0 * wps;
        post_pass<4>(comm + off_v1(a.NC, g.L, S), comm + off_lg(a.NC, g.L, S), wps, nb, S, w2, ncq, bias_s[2], ncq * pc, Q, false, T, stage);
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) generator_pipe_kernel(PipeArgs pa) {
  extern __shared__ __align__(16) float sm[];
  const LatArgs& a = pa.a;
  const int b = blockIdx.x;
  if (b < a.NC) {
    // [NS][LPC+1][32] tagged words | per warp [64 + 128 + 32] floats | [2 NS] ints | causal filter (CTA 0) | [LPC][NS][64] conditioning rows
    volatile u64* xbuf = reinterpret_cast<volatile u64*>(sm);
    float* fbase = sm + 2 * (size_t)pa.NS * (LPC + 1) * 32;
    float* priv = fbase + (threadIdx.x >> 5) * 224;
    int* st_s = reinterpret_cast<int*>(fbase + LPC * 224);
    float* causal_s = fbase + LPC * 224 + 2 * NS_MAX;
    float* pb_s = pa.pb_in_smem ? causal_s + (a.causal_in_smem ? 2 * a.g.Q * C : 0) : nullptr;
    for (int i = threadIdx.x; i < pa.NS * (LPC + 1) * 32; i += THREADS) xbuf[i] = 0ull;
    for (int i = threadIdx.x; i < pa.NS; i += THREADS) {
      st_s[i] = a.g.hdr[i * 4 + 1];
      st_s[pa.NS + i] = a.g.hdr[i * 4 + 0];
    }
    if (b == 0 && a.causal_in_smem)
      for (int i = threadIdx.x; i < 2 * a.g.Q * C; i += THREADS) causal_s[i] = a.g.causal[i];
    __syncthreads();
    chain_warp_ms(pa, b, threadIdx.x >> 5, xbuf, priv, causal_s, pb_s, st_s);
  } else if (b < a.NC + NSAMP) {
    sampler_cta(a, sm, b - a.NC, NSAMP, pa.NS, pa.wps);
  } else if (b < a.NC + NSAMP + a.NP) {
    post_cta_ms(pa, b - a.NC - NSAMP, sm);
  }
}

size_t chain_smem_ms(const GenArgs& g, int NS, int causal_in_smem, int pb_in_smem) {
  return sizeof(float) * (2 * (size_t)NS * (LPC + 1) * 32 + LPC * 224 + 2 * NS_MAX + (causal_in_smem ? 2 * g.Q * C : 0) +
                          (pb_in_smem ? (size_t)LPC * NS * 64 : 0));
}

size_t post_smem_ms(const GenArgs& g, int NP) {
  const int per_group = NP / NG, ncq = (g.Q + per_group - 1) / per_group;
  const size_t LD = (size_t)g.L * C, mx = LD > (size_t)g.S ? LD : (size_t)g.S;
  return sizeof(float) * (LD * PCW + (size_t)g.S * PCW + (size_t)g.S * ncq + SB * mx);
}
size_t chain_smem(const GenArgs& g, int causal_in_smem) {
  return sizeof(float) * (2 * (LPC + 1) * 32 + LPC * 64 + (causal_in_smem ? 2 * g.Q * C : 0));
}
size_t post_smem(const GenArgs& g) { return sizeof(float4) * ((size_t)g.L * C + 2 * g.S + 64); }
size_t sampler_smem(const GenArgs& g) { return (sizeof(double) + sizeof(float)) * (size_t)g.Q; }
}  // namespace

static long long* g_gen_timeline = nullptr;
void set_gen_timeline(long long* p) { g_gen_timeline = p; }

static int comm_words_per_stream(int L, int S, int Q) {
  const int NC = (L + LPC - 1) / LPC;
  return (off_misc(NC, L, S, Q) + 8 + 15) / 16 * 16;
}
int64_t gen_lat_comm_bytes(const wn_config* cfg) {
  return 8LL * NS_MAX * comm_words_per_stream(cfg->n_layers, cfg->skip_channels, cfg->quantization_channels);
}
int gen_pipe_max_streams() { return NS_MAX; }

bool gen_lat_eligible(const GenArgs& a) {
  if (a.streams != 1 || !a.commit || a.C != C || a.n_steps < 1 || a.n_steps >= (1 << 20)) return false;
  if (!a.forced && !a.uniforms && a.n_steps != 1) return false;
  const int NC = (a.L + LPC - 1) / LPC;
  const int NP = (a.S + CS - 1) / CS;
  if ((a.Q + NP - 1) / NP > 4) return false;
  if (a.L * C > 8 * THREADS || a.S > 4 * THREADS || a.Q > 4 * THREADS) return false;      // wait_batch<8> / <4>
  if (NC + 1 + NP > sm_count()) return false;
  if (post_smem(a) > 200 * 1024 || sampler_smem(a) > 200 * 1024) return false;
  return true;
}

// pipelined form: 2 .. NS_MAX streams, every stream advanced by the same number of steps
bool gen_pipe_eligible(const GenArgs& a) {
  if (a.streams < 2 || a.streams > NS_MAX || !a.commit || a.C != C || a.n_steps < 1 || a.n_steps >= (1 << 20)) return false;
  if (!a.forced && !a.uniforms && a.n_steps != 1) return false;
  const int NC = (a.L + LPC - 1) / LPC;
  const int NP = (a.S + CS - 1) / CS;
  if (NP % NG || (NP / NG) * PCW < a.S || ((a.Q + NP / NG - 1) / (NP / NG)) > PCW) return false;      // 16 columns per post CTA, NG groups
  if (a.L * C > 8 * THREADS || a.S > 4 * THREADS || a.Q > 4 * THREADS) return false;      // wait_batch<8> / <4>
  if (NC + NSAMP + NP > sm_count()) return false;
  const int causal_in_smem = (2 * a.Q * C * sizeof(float) <= 64 * 1024) ? 1 : 0;
  const int pb = (a.G > 0 && a.gc_ids) ? 1 : 0;
  if (chain_smem_ms(a, a.streams, causal_in_smem, pb) > 220 * 1024) return false;
  if (post_smem_ms(a, NP) > 220 * 1024 || sampler_smem(a) > 200 * 1024) return false;
  return true;
}

int gen_pipe_run(const GenArgs& g, void* comm, uint32_t launch_seq, cudaStream_t st) {
  if (!gen_pipe_eligible(g) || !comm) return -2;
  PipeArgs pa;
  LatArgs& a = pa.a;
  a.g = g;
  a.comm = (u64*)comm;
  a.tag_base = (launch_seq & 0xFFFu) << 20;
  a.NC = (g.L + LPC - 1) / LPC;
  a.NP = (g.S + CS - 1) / CS;
  a.cq = (g.Q + a.NP - 1) / a.NP;
  a.causal_in_smem = (2 * g.Q * C * sizeof(float) <= 64 * 1024) ? 1 : 0;
  a.timeline = nullptr;
  pa.NS = g.streams;
  pa.wps = comm_words_per_stream(g.L, g.S, g.Q);
  pa.pb_in_smem = (g.G > 0 && g.gc_ids) ? 1 : 0;
  size_t smem = chain_smem_ms(g, pa.NS, a.causal_in_smem, pa.pb_in_smem);
  if (post_smem_ms(g, a.NP) > smem) smem = post_smem_ms(g, a.NP);
  if (sampler_smem(g) > smem) smem = sampler_smem(g);
  cudaError_t e = cudaFuncSetAttribute(generator_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // every CTA spins on words other CTAs produce: all of them must be resident -> cooperative launch
  void* args[] = {&pa};
  e = cudaLaunchCooperativeKernel((const void*)generator_pipe_kernel, dim3(a.NC + NSAMP + a.NP), dim3(THREADS), args, smem, st);
  if (e != cudaSuccess) return (int)e;
  return 0;
}

int gen_lat_run(const GenArgs& g, void* comm, uint32_t launch_seq, cudaStream_t st) {
  if (!gen_lat_eligible(g) || !comm) return -2;
  LatArgs a;
  a.g = g;
  a.comm = (u64*)comm;
  a.tag_base = (launch_seq & 0xFFFu) << 20;
  a.NC = (g.L + LPC - 1) / LPC;
  a.NP = (g.S + CS - 1) / CS;
  a.cq = (g.Q + a.NP - 1) / a.NP;
  a.causal_in_smem = (2 * g.Q * C * sizeof(float) <= 128 * 1024) ? 1 : 0;
  a.timeline = g_gen_timeline;
  size_t smem = chain_smem(g, a.causal_in_smem);
  if (post_smem(g) > smem) smem = post_smem(g);
  if (sampler_smem(g) > smem) smem = sampler_smem(g);
  cudaError_t e = cudaFuncSetAttribute(generator_lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // every CTA spins on words other CTAs produce: all of them must be resident -> cooperative launch
  void* args[] = {&a};
  e = cudaLaunchCooperativeKernel((const void*)generator_lat_kernel, dim3(a.NC + 1 + a.NP), dim3(THREADS), args, smem, st);
  if (e != cudaSuccess) return (int)e;
  return 0;
}

}  // namespace wn
