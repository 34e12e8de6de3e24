// postprocess2 (the last 1x1 convolution) fused with the softmax cross entropy of the training step:
//     logits[m][:] = A2[m][:] . W2 + b2                 model.py:438-440
//     loss         = mean_m  -log softmax(logits[m])[id[m+1]]        model.py:654-666 (rows t = T-1 carry no label)
//     dlogits[m][:] = (softmax(logits[m]) - onehot) * scale          (TF's xent kernel: also for the label-less rows)
// With Q = 256 one 128 x 256 accumulator holds 128 complete logits rows in TMEM, so the logits never travel to HBM: the
// epilogue reads them twice out of TMEM (row maximum / sum, then the gradient), writes the gradient as fp16 (the operand
// of the two GEMMs that consume it) and sums its columns (the bias gradient) on the way out.  What the unfused path does
// with one GEMM launch, a 102 MB logits round trip, the cross-entropy kernel and a column-sum launch.
//
// Structure (persistent CTAs, one per SM, 128-row tiles):
//   warp 0      TMA producer: A [128][64 halfs] + W2 [256][64 halfs] per stage, 3 stages
//   warp 1      MMA issuer: 4 x tcgen05.mma.kind::f16 (M 128, N 256, K 16) per stage into one of two TMEM accumulators
//   warps 2-9   epilogue: two warps per TMEM lane quadrant, 128 columns each; the halves of a row meet through shared memory
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "split_common.cuh"

namespace wn {

namespace {

constexpr int PX_Q = 256;
constexpr int PX_STAGES = 3;
constexpr uint32_t PX_A = 128 * 128;            // [128 rows][64 halfs]
constexpr uint32_t PX_B = PX_Q * 128;           // [256 rows][64 halfs]
constexpr uint32_t PX_STAGE = PX_A + PX_B;
constexpr uint32_t PX_OUT = 32 * 128;           // one [32 rows][64 halfs] output block
constexpr int PX_THREADS = 320;

struct PxArgs {
  const float* bias;          // [256] or null
  const int32_t* ids;         // [M] class ids; the label of row m is ids[m + 1]
  float* partials;            // [gridDim.x] loss sums
  float* colsum;              // [256] += column sums of the gradient * colsum_scale (null: not wanted)
  int M, T, K;
  float scale16;              // gradient scale of the fp16 copy
  float colsum_scale;
};

static int make_map_h2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}

// two floats -> packed fp16 pair (lo = the lower half), saturating
__device__ __forceinline__ uint32_t pack_sat_h(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(PX_THREADS, 1)
post2_xent_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapG, const __grid_constant__ PxArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* out_stage = smem + PX_STAGES * PX_STAGE;      // [8 warps][2][32 rows][64 halfs]
  __shared__ __align__(8) uint64_t full_bar[PX_STAGES], empty_bar[PX_STAGES], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[PX_Q];
  __shared__ float colsum_s[PX_Q];
  __shared__ float2 xch[2][2][128];      // [tile parity][column half][row]: (maximum, sum of exp) of half a row
  __shared__ double red_s[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (a.M + 127) / 128;
  const int nk = a.K / 64;

  for (int i = tid; i < PX_Q; i += blockDim.x) {
    bias_s[i] = a.bias ? __ldg(a.bias + i) : 0.f;
    colsum_s[i] = 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < PX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % PX_STAGES;
          mbar_wait(&empty_bar[s], ((it / PX_STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], PX_STAGE);
          unsigned char* sa = smem + s * PX_STAGE;
          tma_load_2d(sa, &mapA, &full_bar[s], i * 64, tile * 128);
          tma_load_2d(sa + PX_A, &mapB, &full_bar[s], i * 64, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID = idesc_f16(128, PX_Q);
      uint32_t it = 0, local = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local) {
        const uint32_t acc = local & 1;
        mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % PX_STAGES;
          mbar_wait(&full_bar[s], (it / PX_STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * PX_STAGE);
          const uint64_t da = kmajor_desc(sa), db = kmajor_desc(sa + PX_A);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16_ss(tmem + acc * PX_Q, da + 2 * k, db + 2 * k, ID, (i > 0 || k > 0) ? 1u : 0u);
          mma_commit(&empty_bar[s]);
        }
        mma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    const int r = quad * 32 + lane;      // row of the tile = TMEM lane
    const int cbase = half * 128;
    unsigned char* my_out = out_stage + ew * 2 * PX_OUT;
    constexpr float LOG2E = 1.4426950408889634f;
    float colacc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    double loss_local = 0.0;
    uint32_t local = 0, out_cnt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local) {
      const uint32_t acc = local & 1;
      const int m = tile * 128 + r;
      const bool row_ok = m < a.M;
      int target = -1;
      if (row_ok && (m % a.T) < a.T - 1) {
        target = __ldg(a.ids + m + 1);
        if (target < 0 || target >= PX_Q) target = -1;
      }
      mbar_wait(&tmem_full[acc], (local >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem + acc * PX_Q + ((uint32_t)(quad * 32) << 16) + cbase;
      // ---- pass 1: maximum and sum of exponentials of this half of the row (online) ----
      float mx = -INFINITY, sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tbase + c0, v);
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(v[j]) + bias_s[cbase + c0 + j];
          v[j] = __float_as_uint(x);
          cm = fmaxf(cm, x);
        }
        const float nm = fmaxf(mx, cm);
        float cs = 0.f;
        const float off = nm * LOG2E;
#pragma unroll
        for (int j = 0; j < 32; ++j) cs += ex2(fmaf(__uint_as_float(v[j]), LOG2E, -off));
        sum = sum * ex2((mx - nm) * LOG2E) + cs;
        mx = nm;
      }
      xch[acc][half][r] = make_float2(mx, sum);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 other = xch[acc][half ^ 1][r];
      const float gm = fmaxf(mx, other.x);
      const float gs = sum * ex2((mx - gm) * LOG2E) + other.y * ex2((other.x - gm) * LOG2E);
      const float inv = a.scale16 / gs;
      const float goff = gm * LOG2E;
      // ---- pass 2: gradient rows -> fp16 staging blocks of 64 columns -> TMA stores; column sums from the blocks ----
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        const uint32_t buf = out_cnt & 1u;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the store issued two blocks ago has read its block
        __syncwarp();
        unsigned char* os = my_out + buf * PX_OUT + lane * 128;
#pragma unroll 1
        for (int hc = 0; hc < 2; ++hc) {
          const int c0 = blk * 64 + hc * 32;
          uint32_t v[32];
          tmem_ld32(tbase + c0, v);
          float g[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[j]) + bias_s[cbase + c0 + j];
            g[j] = ex2(fmaf(x, LOG2E, -goff)) * inv;
          }
          const int tj = target - (cbase + c0);
          if (tj >= 0 && tj < 32) {
            float xt = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j == tj) { xt = __uint_as_float(v[j]) + bias_s[cbase + c0 + j]; g[j] -= a.scale16; }
            loss_local += (double)((gm + logf(gs)) - xt);
          }
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] = 0.f;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w;
            w.x = pack_sat_h(g[8 * q], g[8 * q + 1]); w.y = pack_sat_h(g[8 * q + 2], g[8 * q + 3]);
            w.z = pack_sat_h(g[8 * q + 4], g[8 * q + 5]); w.w = pack_sat_h(g[8 * q + 6], g[8 * q + 7]);
            *reinterpret_cast<uint4*>(os + ((uint32_t)((hc * 4 + q) ^ (lane & 7)) << 4)) = w;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&mapG), "r"(smem_u32(my_out + buf * PX_OUT)), "r"(cbase + blk * 64), "r"(tile * 128 + quad * 32) : "memory");
          bulk_commit();
        }
        if (a.colsum) {      // lane j: columns 2j, 2j+1 of the block, over its 32 rows
          const unsigned char* ob = my_out + buf * PX_OUT;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const __half2 h = *reinterpret_cast<const __half2*>(ob + rr * 128 + ((uint32_t)((lane >> 2) ^ (rr & 7)) << 4) + (lane & 3) * 4);
            const float2 f = __half22float2(h);
            s0 += f.x; s1 += f.y;
          }
          colacc[blk][0] += s0; colacc[blk][1] += s1;
        }
        ++out_cnt;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (lane == 0) bulk_wait0();
    // ---- loss and column sums of this CTA ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, o);
    if (lane == 0) red_s[ew] = loss_local;
    if (a.colsum) {
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        atomicAdd(&colsum_s[cbase + blk * 64 + 2 * lane], colacc[blk][0]);
        atomicAdd(&colsum_s[cbase + blk * 64 + 2 * lane + 1], colacc[blk][1]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int etid = tid - 64;
    if (a.colsum) atomicAdd(a.colsum + etid, colsum_s[etid] * a.colsum_scale);
    if (etid == 0) {
      double s = 0.0;
      for (int i = 0; i < 8; ++i) s += red_s[i];
      a.partials[blockIdx.x] = (float)s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

// A16: [M][K] fp16 (postprocess2's input), W16: [256][K] fp16 (K-major weight copy), g16: [M][256] fp16 gradient out
// (scaled by scale16), loss_out = loss_scale * sum of the row losses; bias_grad (nullable, [256]) += colsum_scale * column
// sums of g16.  partials: >= sm_count() floats.  Returns -2 when the shape is not the fused one (Q != 256, K % 64).
int post2_xent(const void* A16, int lda, const void* W16, int ldw, const float* bias, const int32_t* ids, int M, int T, int K,
               int Q, float loss_scale, float* partials, float* loss_out, void* g16, float scale16, float* bias_grad,
               float colsum_scale, cudaStream_t st) {
  if (Q != PX_Q || K < 64 || (K % 64) || M < 1) return -2;
  CUtensorMap mA, mB, mG;
  int rc = make_map_h2d(&mA, A16, M, K, lda, 128);
  if (rc) return rc;
  rc = make_map_h2d(&mB, W16, Q, K, ldw, PX_Q);
  if (rc) return rc;
  rc = make_map_h2d(&mG, g16, M, Q, Q, 32);
  if (rc) return rc;
  PxArgs a;
  a.bias = bias; a.ids = ids; a.partials = partials; a.colsum = bias_grad;
  a.M = M; a.T = T; a.K = K; a.scale16 = scale16; a.colsum_scale = colsum_scale;
  const size_t smem = 1024 + PX_STAGES * PX_STAGE + 8 * 2 * PX_OUT;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(post2_xent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -5;
    attr = true;
  }
  const int n_tiles = (M + 127) / 128;
  int grid = sm_count();
  if (grid > n_tiles) grid = n_tiles;
  post2_xent_kernel<<<grid, PX_THREADS, smem, st>>>(mA, mB, mG, a);
  WN_CHECK_LAUNCH();
  return xent_finalize(partials, grid, loss_scale, loss_out, st);
}

}  // namespace wn
