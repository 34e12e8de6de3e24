// Gated residual block on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), C = 32 channels.
// Reference: wavenet/model.py:236-330 (_create_dilation_layer) with the two ops.causal_conv calls
// (wavenet/ops.py:46-62) and, for the backward kernels, TF autodiff of the same lines.
//
// One CTA processes tiles of 128 consecutive time steps of one batch element:
//   * the activation tile x[t0 : t0+128] and the dilated "past" tile x[t0-d : t0-d+128] arrive by
//     TMA from a 3-D tensor map [B][T][32]; time coordinates outside [0,T) are zero filled, which IS
//     the causal padding -- no pad / time_to_batch / transpose / batch_to_time tensors exist;
//   * a row of 32 fp32 channels is exactly one 128-byte swizzle row, i.e. the K-major A operand of
//     tcgen05.mma.kind::tf32 (M = 128 time steps);
//   * accumulators live in TMEM, thread r of the CTA owns time step t0 + r (TMEM lane r);
//   * intermediate operands (z, lo parts, gradients) are written back to shared memory in the same
//     swizzled layout by the threads (fence.proxy.async) and fed to the next MMA: nothing but x, z and
//     x' touches HBM.
// Forward products are split-precision (hi + lo TF32 terms, 3 MMAs per product): DESIGN.md section 6.
#include "common.cuh"
#include "kernels.h"
#include "umma_common.cuh"

namespace wn {
using namespace umma;

namespace {
constexpr int C = 32;
constexpr int TM = 128;                 // time steps per tile
constexpr uint32_t TILE = TM * 128;     // bytes of a [128][32] fp32 tile

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// stage a weight matrix as a K-major B operand: rows n (output channel), cols k (input channel), hi/lo split
template <typename F>
__device__ __forceinline__ void stage_b(unsigned char* hi, unsigned char* lo, int n_rows, F&& w_of) {
  for (int i = threadIdx.x; i < n_rows * 32; i += blockDim.x) {
    const int n = i >> 5, k = i & 31;
    const float w = w_of(n, k);
    const float h = round_tf32(w);
    *reinterpret_cast<float*>(hi + swz(n, k)) = h;
    if (lo) *reinterpret_cast<float*>(lo + swz(n, k)) = round_tf32(w - h);
  }
}
}  // namespace

struct FwdArgs {
  float* xout;
  float* zc; int ldz;          // Zcat + l*C, row pitch ldz
  float* zcT; int ldm;         // ZcatT + l*C*ldm (nullable): transposed copy for the skip weight-gradient GEMM
  const float *wf, *wg, *dense, *prebias, *dense_bias;
  int B, T, d, is_last;
};

__global__ void __launch_bounds__(128, 2)
block_fwd_umma_kernel(const __grid_constant__ CUtensorMap mapX, FwdArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* Xc = smem;
  unsigned char* Xp = smem + TILE;
  unsigned char* L0 = smem + 2 * TILE;     // lo(x_past), later hi(z)
  unsigned char* L1 = smem + 3 * TILE;     // lo(x_cur),  later lo(z)
  unsigned char* W0h = smem + 4 * TILE;    // [64][32] past-tap weights (filter | gate), hi
  unsigned char* W0l = W0h + 8192;
  unsigned char* W1h = W0l + 8192;         // current tap
  unsigned char* W1l = W1h + 8192;
  unsigned char* Wdh = W1l + 8192;         // [32][32] dense^T
  unsigned char* Wdl = Wdh + 4096;
  __shared__ __align__(8) uint64_t bar_tma, bar_m1, bar_m2;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_m1, 1);
    mbar_init(&bar_m2, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  stage_b(W0h, W0l, 64, [&](int n, int k) { return n < C ? a.wf[k * C + n] : a.wg[k * C + (n - C)]; });
  stage_b(W1h, W1l, 64, [&](int n, int k) { return n < C ? a.wf[(C + k) * C + n] : a.wg[(C + k) * C + (n - C)]; });
  if (!a.is_last) stage_b(Wdh, Wdl, 32, [&](int n, int k) { return a.dense[k * C + n]; });
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t ID64 = idesc_tf32(128, 64), ID32 = idesc_tf32(128, 32);
  const uint64_t dXc = kmajor_desc(smem_u32(Xc)), dXp = kmajor_desc(smem_u32(Xp));
  const uint64_t dL0 = kmajor_desc(smem_u32(L0)), dL1 = kmajor_desc(smem_u32(L1));
  const uint64_t dW0h = kmajor_desc(smem_u32(W0h)), dW0l = kmajor_desc(smem_u32(W0l));
  const uint64_t dW1h = kmajor_desc(smem_u32(W1h)), dW1l = kmajor_desc(smem_u32(W1l));
  const uint64_t dWdh = kmajor_desc(smem_u32(Wdh)), dWdl = kmajor_desc(smem_u32(Wdl));

  const int n_tt = (a.T + TM - 1) / TM;
  const int n_tiles = a.B * n_tt;
  const int r = tid;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    const uint32_t par = it & 1;
    if (tid == 0) {
      mbar_expect_tx(&bar_tma, 2 * TILE);
      tma_load_3d(Xc, &mapX, &bar_tma, 0, t0, b);
      tma_load_3d(Xp, &mapX, &bar_tma, 0, t0 - a.d, b);
    }
    mbar_wait(&bar_tma, par);
    if (tid == 0) {   // hi*hi terms: the tensor core reads the upper 19 bits of the raw fp32 tile
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXp + 2 * k, dW0h + 2 * k, ID64, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXc + 2 * k, dW1h + 2 * k, ID64, 1);
    }
    // lo parts of this thread's rows: x - trunc_tf32(x), same swizzled position
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t off = (uint32_t)r * 128 + ((uint32_t)(j ^ (r & 7)) << 4);
      float4 v = *reinterpret_cast<const float4*>(Xc + off);
      *reinterpret_cast<float4*>(L1 + off) =
          make_float4(v.x - trunc_tf32(v.x), v.y - trunc_tf32(v.y), v.z - trunc_tf32(v.z), v.w - trunc_tf32(v.w));
      v = *reinterpret_cast<const float4*>(Xp + off);
      *reinterpret_cast<float4*>(L0 + off) =
          make_float4(v.x - trunc_tf32(v.x), v.y - trunc_tf32(v.y), v.z - trunc_tf32(v.z), v.w - trunc_tf32(v.w));
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dL0 + 2 * k, dW0h + 2 * k, ID64, 1);   // lo(x_past) . hi(W0)
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXp + 2 * k, dW0l + 2 * k, ID64, 1);   // hi(x_past) . lo(W0)
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dL1 + 2 * k, dW1h + 2 * k, ID64, 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXc + 2 * k, dW1l + 2 * k, ID64, 1);
      mma_commit(&bar_m1);
    }
    mbar_wait(&bar_m1, par);
    tc_fence_after();

    const bool valid = (t0 + r) < a.T;
    const size_t m = (size_t)b * a.T + t0 + r;
    float z[32];
    {
      uint32_t fv[32], gv[32];
      tmem_ld32(lane_addr + 0, fv);
      tmem_ld32(lane_addr + 32, gv);
      const float* pb = a.prebias + (size_t)b * 64;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        z[j] = tanh_f(__uint_as_float(fv[j]) + __ldg(pb + j)) * sigmoid_f(__uint_as_float(gv[j]) + __ldg(pb + 32 + j));
    }
    // z -> Zcat (tf32-rounded: it feeds the single-pass skip GEMM), its transposed copy, and the
    // hi/lo A operand of the dense product
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = round_tf32(z[4 * j + e]);
        l[e] = round_tf32(z[4 * j + e] - h[e]);
      }
      if (valid) *reinterpret_cast<float4*>(a.zc + m * a.ldz + 4 * j) = make_float4(h[0], h[1], h[2], h[3]);
      if (a.zcT && valid) {
#pragma unroll
        for (int e = 0; e < 4; ++e) a.zcT[(size_t)(4 * j + e) * a.ldm + m] = h[e];
      }
      if (!a.is_last) {
        const uint32_t off = (uint32_t)r * 128 + ((uint32_t)(j ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(L0 + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(L1 + off) = make_float4(l[0], l[1], l[2], l[3]);
      }
    }
    if (!a.is_last) {
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem + 64, dL0 + 2 * k, dWdh + 2 * k, ID32, k > 0);   // hi(z) . hi(Wd)
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem + 64, dL1 + 2 * k, dWdh + 2 * k, ID32, 1);       // lo(z) . hi(Wd)
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem + 64, dL0 + 2 * k, dWdl + 2 * k, ID32, 1);       // hi(z) . lo(Wd)
        mma_commit(&bar_m2);
      }
      mbar_wait(&bar_m2, par);
      tc_fence_after();
      uint32_t ov[32];
      tmem_ld32(lane_addr + 64, ov);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t off = (uint32_t)r * 128 + ((uint32_t)(j ^ (r & 7)) << 4);
        const float4 xv = *reinterpret_cast<const float4*>(Xc + off);
        float4 bd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.dense_bias) bd = __ldg(reinterpret_cast<const float4*>(a.dense_bias + 4 * j));
        if (valid)
          *reinterpret_cast<float4*>(a.xout + m * C + 4 * j) =
              make_float4(xv.x + __uint_as_float(ov[4 * j]) + bd.x, xv.y + __uint_as_float(ov[4 * j + 1]) + bd.y,
                          xv.z + __uint_as_float(ov[4 * j + 2]) + bd.z, xv.w + __uint_as_float(ov[4 * j + 3]) + bd.w);
      }
    }
    tc_fence_before();
    __syncthreads();   // shared tiles and TMEM columns are free for the next tile
  }
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int block_fwd_umma(const float* x, float* xout, float* zc, int ldz, float* zcT, int ldm, const float* wf,
                   const float* wg, const float* dense, const float* prebias, const float* dense_bias, int B, int T,
                   int d, int is_last, cudaStream_t st) {
  CUtensorMap mapX;
  int rc = make_map_3d(&mapX, x, B, T, C, C, TM);
  if (rc) return rc;
  FwdArgs a;
  a.xout = xout; a.zc = zc; a.ldz = ldz; a.zcT = zcT; a.ldm = ldm; a.wf = wf; a.wg = wg; a.dense = dense;
  a.prebias = prebias; a.dense_bias = dense_bias; a.B = B; a.T = T; a.d = d; a.is_last = is_last;
  const size_t smem = 1024 + 4 * TILE + 4 * 8192 + 2 * 4096;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(block_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int n_tiles = B * ((T + TM - 1) / TM);
  int grid = n_tiles;
  const int cap = 2 * sm_count();
  if (grid > cap) grid = cap;
  block_fwd_umma_kernel<<<grid, 128, smem, st>>>(mapX, a);
  WN_CHECK_LAUNCH();
  return 0;
}

}  // namespace wn
