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
//     x' touches HBM;
//   * the weights of a layer are pre-arranged once per step into the exact shared-memory image
//     (swizzled K-major B operands, hi/lo split) and arrive with one cp.async.bulk.
// Forward products are split-precision (hi + lo TF32 terms, 3 MMAs per product): DESIGN.md section 6.
//
// Backward of a layer is three kernels:
//   bwd_pre  : recompute pre-activations, dz = dz_skip + dx'.Wd^T, dpre = [df|dg]  -> dpre
//   wgrad    : one GEMM over time (K = T) gives every weight / bias gradient of the layer:
//              [x ; x[t-d] ; z ; 1]^T . [dpre ; dx']  with both operands read MN-major straight from the
//              [time][channel] activations (no transposed copies)
//   bwd_dx   : dx = dx' + dpre[t].Wcur^T + dpre[t+d].Wpast^T                       -> dx
#include <cstdlib>

#include <cuda_fp16.h>

#include "../../include/wavenet_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "umma_common.cuh"

namespace wn {
using namespace umma;

namespace {
constexpr int C = 32;
constexpr int TM = 128;                 // time steps per tile
constexpr uint32_t TILE = TM * 128;     // bytes of a [128][32] fp32 tile
constexpr uint32_t IMG_FWD = 4 * 8192 + 2 * 4096;    // W0h W0l W1h W1l | Wdh Wdl
constexpr uint32_t IMG_PRE = 2 * 8192 + 4096;        // W0h W1h | WdB
constexpr uint32_t IMG_DX = 4 * 4096;                // Bcf Bcg Bpf Bpg
constexpr uint32_t IMG_ALL = IMG_FWD + IMG_PRE + IMG_DX;

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// K-major B operand image: rows n (output channel), cols k (input channel), hi (and lo) split
template <typename F>
__device__ __forceinline__ void stage_b(unsigned char* hi, unsigned char* lo, int n_rows, F&& w_of) {
  for (int i = threadIdx.x; i < n_rows * 32; i += blockDim.x) {
    const int k = i / n_rows, n = i % n_rows;     // n fastest: coalesced reads of [k][n] weights
    const float w = w_of(n, k);
    const float h = round_tf32(w);
    *reinterpret_cast<float*>(hi + swz(n, k)) = h;
    if (lo) *reinterpret_cast<float*>(lo + swz(n, k)) = round_tf32(w - h);
  }
}

// the three weight images of one layer (pointers may be shared or global memory)
__device__ __forceinline__ void build_images(unsigned char* fwd, unsigned char* pre, unsigned char* dx,
                                             const float* __restrict__ wf, const float* __restrict__ wg,
                                             const float* __restrict__ dense, bool have_dense) {
  auto w0 = [&](int n, int k) { return n < C ? wf[k * C + n] : wg[k * C + (n - C)]; };
  auto w1 = [&](int n, int k) { return n < C ? wf[(C + k) * C + n] : wg[(C + k) * C + (n - C)]; };
  if (fwd) {
    stage_b(fwd, fwd + 8192, 64, w0);
    stage_b(fwd + 16384, fwd + 24576, 64, w1);
    if (have_dense) stage_b(fwd + 32768, fwd + 36864, 32, [&](int n, int k) { return dense[k * C + n]; });   // n = r, k = d
  }
  if (pre) {
    stage_b(pre, nullptr, 64, w0);
    stage_b(pre + 8192, nullptr, 64, w1);
    if (have_dense) stage_b(pre + 16384, nullptr, 32, [&](int n, int k) { return dense[n * C + k]; });      // n = d, k = r
  }
  if (dx) {   // n = residual channel r, k = dilation channel d
    stage_b(dx, nullptr, 32, [&](int n, int k) { return wf[(C + n) * C + k]; });
    stage_b(dx + 4096, nullptr, 32, [&](int n, int k) { return wg[(C + n) * C + k]; });
    stage_b(dx + 8192, nullptr, 32, [&](int n, int k) { return wf[n * C + k]; });
    stage_b(dx + 12288, nullptr, 32, [&](int n, int k) { return wg[n * C + k]; });
  }
}
}  // namespace

__global__ void block_images_kernel(unsigned char* __restrict__ img, const float* __restrict__ filter,
                                    const float* __restrict__ gate, const float* __restrict__ dense) {
  const int l = blockIdx.x;
  unsigned char* base = img + (size_t)l * IMG_ALL;
  build_images(base, base + IMG_FWD, base + IMG_FWD + IMG_PRE, filter + (size_t)l * 2 * C * C,
               gate + (size_t)l * 2 * C * C, dense + (size_t)l * C * C, true);
}

int64_t block_images_bytes(int L) { return (int64_t)L * IMG_ALL; }
uint32_t block_img_off_pre() { return IMG_FWD; }
uint32_t block_img_off_dx() { return IMG_FWD + IMG_PRE; }
uint32_t block_img_stride() { return IMG_ALL; }
int block_images(unsigned char* img, const float* filter, const float* gate, const float* dense, int L, cudaStream_t st) {
  block_images_kernel<<<L, 256, 0, st>>>(img, filter, gate, dense);
  WN_CHECK_LAUNCH();
  return 0;
}

// =========================================================================================
// forward
// =========================================================================================
struct FwdArgs {
  float* xout;
  float* zc; int ldz;               // Zcat + l*C, row pitch ldz
  const unsigned char* img;         // IMG_FWD bytes (nullable -> built in the kernel from wf/wg/dense)
  const float *wf, *wg, *dense, *prebias, *dense_bias;
  int B, T, d, is_last;
  long long* timeline;              // debug: per-phase clock64 stamps of CTA 0 (wn_debug_timeline), else null
  int pdl_next;
};

static long long* g_timeline = nullptr;
void set_block_timeline(long long* p) { g_timeline = p; }
#define TL(i)                                                                          \
  do {                                                                                 \
    if (a.timeline && blockIdx.x == 0 && tid == 0 && it < 4) a.timeline[it * 8 + (i)] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(256, 2)
block_fwd_umma_kernel(const __grid_constant__ CUtensorMap mapX, FwdArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the .shared provenance
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
  __shared__ __align__(8) uint64_t bar_tma, bar_m1, bar_m2, bar_w;
  __shared__ uint32_t tmem_slot;
  __shared__ float pb_s[64];
  __shared__ float bd_s[32];

  // 256 threads: thread (r, half) owns time step t0 + r and channels [16*half, 16*half + 16) of every
  // 32-channel quantity of that step (TMEM lane r is reachable from warps r/32 and r/32 + 4).
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;
  const int tl_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;
  if (a.timeline && tid == 0 && tl_slot >= 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
    a.timeline[32 + 3 * tl_slot] = (long long)g;
  }
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_m1, 1);
    mbar_init(&bar_m2, 1);
    mbar_init(&bar_w, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid < 32) bd_s[tid] = a.dense_bias ? a.dense_bias[tid] : 0.f;
  __syncthreads();
  if (a.img) {
    if (tid == 0) {
      mbar_expect_tx(&bar_w, IMG_FWD);
      bulk_g2s(W0h, a.img, IMG_FWD, &bar_w);
    }
  } else {
    build_images(W0h, nullptr, nullptr, a.wf, a.wg, a.dense, !a.is_last);
    fence_async_smem();
  }

  const int n_tt = (a.T + TM - 1) / TM;
  const int n_tiles = a.B * n_tt;
  auto issue_loads = [&](int tile) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    mbar_expect_tx(&bar_tma, 2 * TILE);
    tma_load_3d(Xc, &mapX, &bar_tma, 0, t0, b);
    tma_load_3d(Xp, &mapX, &bar_tma, 0, t0 - a.d, b);
  };
  pdl_wait();      // x (previous layer's output) is complete and visible from here on
  if (a.pdl_next) pdl_trigger();
  if (tid == 0 && (int)blockIdx.x < n_tiles) issue_loads(blockIdx.x);
  if (a.img) mbar_wait(&bar_w, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16 * half;
  constexpr uint32_t ID64 = idesc_tf32(128, 64), ID32 = idesc_tf32(128, 32);
  const uint64_t dXc = kmajor_desc(smem_u32(Xc)), dXp = kmajor_desc(smem_u32(Xp));
  const uint64_t dL0 = kmajor_desc(smem_u32(L0)), dL1 = kmajor_desc(smem_u32(L1));
  const uint64_t dW0h = kmajor_desc(smem_u32(W0h)), dW0l = kmajor_desc(smem_u32(W0l));
  const uint64_t dW1h = kmajor_desc(smem_u32(W1h)), dW1l = kmajor_desc(smem_u32(W1l));
  const uint64_t dWdh = kmajor_desc(smem_u32(Wdh)), dWdl = kmajor_desc(smem_u32(Wdl));

  int it = 0;
  int pb_batch = -1;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    const uint32_t par = it & 1;
    if (b != pb_batch) {   // block-uniform: conditioning + bias row of this batch element
      __syncthreads();
      if (tid < 64) pb_s[tid] = a.prebias[(size_t)b * 64 + tid];
      pb_batch = b;
      __syncthreads();      // (readers sit behind mbarrier waits only: without this a slow writer warp races them)
    }
    TL(0);
    mbar_wait(&bar_tma, par);
    TL(1);
    if (tid == 0) {   // hi(x) terms first (the tensor core reads the upper 19 bits of the raw fp32 tile): they run
      tc_fence_after();   // while the threads split off the lo parts below
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXp + 2 * k, dW0h + 2 * k, ID64, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXc + 2 * k, dW1h + 2 * k, ID64, 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXp + 2 * k, dW0l + 2 * k, ID64, 1);   // hi(x_past) . lo(W0)
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXc + 2 * k, dW1l + 2 * k, ID64, 1);   // hi(x_cur)  . lo(W1)
    }
    // lo parts of this thread's half rows: x - trunc_tf32(x), same swizzled position; x_cur stays in registers
    float4 xr[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * half + jj;
      const uint32_t off = (uint32_t)r * 128 + ((uint32_t)(j ^ (r & 7)) << 4);
      float4 v = *reinterpret_cast<const float4*>(Xc + off);
      xr[jj] = v;
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
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dL1 + 2 * k, dW1h + 2 * k, ID64, 1);   // lo(x_cur)  . hi(W1)
      mma_commit(&bar_m1);
    }
    const bool valid = (t0 + r) < a.T;
    const size_t m = (size_t)b * a.T + t0 + r;
    TL(2);
    mbar_wait(&bar_m1, par);
    TL(3);
    tc_fence_after();
    // both input tiles are consumed: prefetch the next tile of this CTA behind the epilogue
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) issue_loads(tile + gridDim.x);

    float z[16];
    {
      uint32_t fv[16], gv[16];
      tmem_ld16(lane_addr + 0, fv);
      tmem_ld16(lane_addr + 32, gv);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        z[j] = gated_fast(__uint_as_float(fv[j]) + pb_s[16 * half + j], __uint_as_float(gv[j]) + pb_s[32 + 16 * half + j]);
    }
    // z -> Zcat (tf32-rounded: it feeds the single-pass skip GEMM) and the hi/lo A operand of the dense product
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * half + jj;
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = round_tf32(z[4 * jj + e]);
        l[e] = round_tf32(z[4 * jj + e] - h[e]);
      }
      if (valid) *reinterpret_cast<float4*>(a.zc + m * a.ldz + 4 * j) = make_float4(h[0], h[1], h[2], h[3]);
      if (!a.is_last) {
        const uint32_t off = (uint32_t)r * 128 + ((uint32_t)(j ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(L0 + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(L1 + off) = make_float4(l[0], l[1], l[2], l[3]);
      }
    }
    TL(4);
    if (!a.is_last) {
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      TL(5);
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
      TL(6);
      tc_fence_after();
      uint32_t ov[16];
      tmem_ld16(lane_addr + 64, ov);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int c = 16 * half + 4 * jj;
        const float4 xv = xr[jj];
        if (valid)
          *reinterpret_cast<float4*>(a.xout + m * C + c) =
              make_float4(xv.x + __uint_as_float(ov[4 * jj]) + bd_s[c], xv.y + __uint_as_float(ov[4 * jj + 1]) + bd_s[c + 1],
                          xv.z + __uint_as_float(ov[4 * jj + 2]) + bd_s[c + 2], xv.w + __uint_as_float(ov[4 * jj + 3]) + bd_s[c + 3]);
      }
    }
    tc_fence_before();
    __syncthreads();   // shared tiles and TMEM columns are free for the next tile
    TL(7);
    if (a.timeline && tid == 0 && tl_slot >= 0 && it == 0) {
      unsigned long long g;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
      a.timeline[32 + 3 * tl_slot + 1] = (long long)g;   // end of the first tile
    }
  }
  if (a.timeline && tid == 0 && tl_slot >= 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
    a.timeline[32 + 3 * tl_slot + 2] = (long long)g;
  }
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int block_fwd_umma(const float* x, float* xout, float* zc, int ldz, const unsigned char* img, const float* wf, const float* wg, const float* dense,
                   const float* prebias, const float* dense_bias, int B, int T, int d, int is_last, cudaStream_t st) {
  CUtensorMap mapX;
  int rc = make_map_3d(&mapX, x, B, T, C, C, TM);
  if (rc) return rc;
  FwdArgs a;
  a.xout = xout; a.zc = zc; a.ldz = ldz; a.img = img; a.wf = wf; a.wg = wg;
  a.dense = dense; a.prebias = prebias; a.dense_bias = dense_bias; a.B = B; a.T = T; a.d = d; a.is_last = is_last;
  a.timeline = g_timeline;
  a.pdl_next = 0;      // callers mix this kernel with plain launches (stand-alone C ABI entry, WN_BLOCK_FWD=tf32)
  const size_t smem = 1024 + 4 * TILE + IMG_FWD;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(block_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int n_tiles = B * ((T + TM - 1) / TM);
  int grid = n_tiles;
  const int cap = 2 * sm_count();
  if (grid > cap) grid = cap;
  { cudaError_t e = launch_pdl(block_fwd_umma_kernel, dim3(grid), dim3(256), smem, st, mapX, a); if (e != cudaSuccess) return (int)e; }
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_FWD);
  return 0;
}

// =========================================================================================
// backward 1/3: dpre = [df | dg]
// =========================================================================================
struct PreArgs {
  float* dpre;                 // [M][64]
  const unsigned char* img;    // IMG_PRE bytes
  const float* prebias;
  int B, T, d, is_last, zcol;  // zcol: column of this layer inside dZcat
  int pdl_next;                // the next kernel in the stream is launched programmatically and waits (common.cuh)
  float dz_scale;              // DZ16: the skip-path gradient arrives as fp16 in a domain scaled by 1 / dz_scale
};

// DZ16 = false: the skip-path gradient is an fp32 tile and the two dpre halves are staged in the (dead) input tiles --
//   the next tile's Dz / Dn loads have to wait until the stores have READ those tiles (~1500 cycles) and then take their
//   own ~1500 cycles: a serial chain in every tile.
// DZ16 = true (fp16 gradient chain): the fp16 tile is 8 KB, which leaves room (2 CTAs per SM) for separate staging tiles;
//   every input tile of the next tile is requested at the mid-tile barrier and the stores drain on their own.
template <bool DZ16>
__global__ void __launch_bounds__(256, 2)
block_bwd_pre_umma_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDn,
                          const __grid_constant__ CUtensorMap mapDz, const __grid_constant__ CUtensorMap mapDp, PreArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Xc = smem;
  unsigned char* Xp = smem + TILE;
  unsigned char* Dn = smem + 2 * TILE;     // dx' tile (A operand of dx'.Wd^T); !DZ16: later staging of dg
  unsigned char* Dz = smem + 3 * TILE;     // !DZ16: skip-path gradient tile (read by the threads), later staging of df; DZ16: staging of df
  unsigned char* W0 = smem + 4 * TILE;
  unsigned char* W1 = W0 + 8192;
  unsigned char* Wd = W1 + 8192;
  unsigned char* Sg = W0 + IMG_PRE;        // DZ16: staging of dg
  unsigned char* Zq = Sg + TILE;           // DZ16: skip-path gradient tile, [128 rows][32 halfs], 64B swizzle
  unsigned char* Gs = DZ16 ? Sg : Dn;      // where dg is staged
  __shared__ __align__(8) uint64_t bar_x, bar_d, bar_m1, bar_w;
  __shared__ uint32_t tmem_slot;
  __shared__ float pb_s[64];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;   // thread = (time step, 16-channel half), see block_fwd_umma_kernel
  if (tid == 0) {
    mbar_init(&bar_x, 1);
    mbar_init(&bar_d, 1);
    mbar_init(&bar_m1, 1);
    mbar_init(&bar_w, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  __syncthreads();
  const int n_tt = (a.T + TM - 1) / TM;
  const int n_tiles = a.B * n_tt;
  // Two groups of input tiles: x tiles are consumed by the MMAs only and are prefetched as soon as those are
  // done; the Dn / Dz tiles double as the staging tiles of the dpre output (TMA store) and are reloaded late.
  auto issue_x = [&](int tile) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    mbar_expect_tx(&bar_x, 2 * TILE);
    tma_load_3d(Xc, &mapX, &bar_x, 0, t0, b);
    tma_load_3d(Xp, &mapX, &bar_x, 0, t0 - a.d, b);
  };
  // (two halves: the skip-path gradient was written long before this launch, dx' by the direct predecessor)
  auto issue_dz = [&](int tile) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    mbar_expect_tx(&bar_d, (DZ16 ? TM * 64 : TILE) + (a.is_last ? 0 : TILE));
    tma_load_3d(DZ16 ? Zq : Dz, &mapDz, &bar_d, a.zcol, t0, b);
  };
  auto issue_dn = [&](int tile) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    if (!a.is_last) tma_load_3d(Dn, &mapDn, &bar_d, 0, t0, b);
  };
  auto issue_d = [&](int tile) {
    issue_dz(tile);
    issue_dn(tile);
  };
  if (tid == 0) {
    mbar_expect_tx(&bar_w, IMG_PRE);
    bulk_g2s(W0, a.img, IMG_PRE, &bar_w);
  }
  // Everything the first tile reads that is older than the direct predecessor (layer inputs from the forward pass, the
  // skip-path gradient from the GEMM chain) is requested BEFORE the dependency wait (invariant in common.cuh); only dx'
  // -- written by the dx kernel right before this launch -- has to wait.
  if (tid == 0 && (int)blockIdx.x < n_tiles) {
    issue_x(blockIdx.x);
    issue_dz(blockIdx.x);
  }
  pdl_wait();
  if (a.pdl_next) pdl_trigger();
  if (tid == 0 && (int)blockIdx.x < n_tiles) issue_dn(blockIdx.x);
  mbar_wait(&bar_w, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16 * half;
  constexpr uint32_t ID64 = idesc_tf32(128, 64), ID32 = idesc_tf32(128, 32);
  const uint64_t dXc = kmajor_desc(smem_u32(Xc)), dXp = kmajor_desc(smem_u32(Xp)), dDn = kmajor_desc(smem_u32(Dn));
  const uint64_t dW0 = kmajor_desc(smem_u32(W0)), dW1 = kmajor_desc(smem_u32(W1)), dWd = kmajor_desc(smem_u32(Wd));
  const uint32_t row_off = (uint32_t)r * 128;

  int it = 0, pb_batch = -1;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    const uint32_t par = it & 1;
    if (b != pb_batch) {
      __syncthreads();
      if (tid < 64) pb_s[tid] = a.prebias[(size_t)b * 64 + tid];
      pb_batch = b;
      __syncthreads();      // (readers sit behind mbarrier waits only: without this a slow writer warp races them)
    }
    mbar_wait(&bar_x, par);
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXp + 2 * k, dW0 + 2 * k, ID64, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dXc + 2 * k, dW1 + 2 * k, ID64, 1);
    }
    mbar_wait(&bar_d, par);
    if (tid == 0) {
      if (!a.is_last) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem + 64, dDn + 2 * k, dWd + 2 * k, ID32, k > 0);   // dx' . Wd^T
      }
      mma_commit(&bar_m1);
    }
    // gradient coming from the skip path (this thread's half row of the dZcat tile)
    float dz[16];
    if (DZ16) {
      const unsigned char* zr = Zq + (uint32_t)r * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(zr + ((uint32_t)((2 * half + c) ^ ((r >> 1) & 3)) << 4));
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __half22float2(h[q]);
          dz[8 * c + 2 * q] = f.x * a.dz_scale; dz[8 * c + 2 * q + 1] = f.y * a.dz_scale;
        }
      }
    } else {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 v = *reinterpret_cast<const float4*>(Dz + row_off + ((uint32_t)((4 * half + jj) ^ (r & 7)) << 4));
        dz[4 * jj] = v.x; dz[4 * jj + 1] = v.y; dz[4 * jj + 2] = v.z; dz[4 * jj + 3] = v.w;
      }
    }
    mbar_wait(&bar_m1, par);
    tc_fence_after();
    if (DZ16 && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile's stores have left the staging tiles
    // Re-arming an mbarrier while a slow thread has not yet observed the phase it waits for lets the barrier wrap
    // around to the same parity: that thread then waits forever (rare, box dependent hangs).  Every thread must
    // be past its bar_x wait first.
    __syncthreads();
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) {
      issue_x(tile + gridDim.x);                 // x tiles: only the MMAs read them
      if (DZ16) issue_d(tile + gridDim.x);       // dx' tile: read by the finished MMAs; fp16 dz tile: every thread has its values
    }

    const bool valid = (t0 + r) < a.T;
    if (!a.is_last) {
      uint32_t av[16];
      tmem_ld16(lane_addr + 64, av);
#pragma unroll
      for (int j = 0; j < 16; ++j) dz[j] += __uint_as_float(av[j]);
    }
    uint32_t fv[16], gv[16];
    tmem_ld16(lane_addr + 0, fv);
    tmem_ld16(lane_addr + 32, gv);
    float df[16], dg[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float tf, sg;
      gated_parts_fast(__uint_as_float(fv[j]) + pb_s[16 * half + j], __uint_as_float(gv[j]) + pb_s[32 + 16 * half + j], tf, sg);
      const float dzv = valid ? dz[j] : 0.f;
      df[j] = round_tf32(dzv * sg * (1.f - tf * tf));
      dg[j] = round_tf32(dzv * tf * sg * (1.f - sg));
    }
    // dpre = [df | dg] leaves through two TMA stores: df is staged in the Dz tile (each thread overwrites exactly
    // the chunks it read), dg in the Dn tile (read by the finished MMAs only)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const uint32_t off = row_off + ((uint32_t)((4 * half + jj) ^ (r & 7)) << 4);
      *reinterpret_cast<float4*>(Dz + off) = make_float4(df[4 * jj], df[4 * jj + 1], df[4 * jj + 2], df[4 * jj + 3]);
      *reinterpret_cast<float4*>(Gs + off) = make_float4(dg[4 * jj], dg[4 * jj + 1], dg[4 * jj + 2], dg[4 * jj + 3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(&mapDp), "r"(smem_u32(Dz)), "r"(0), "r"(t0), "r"(b) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(&mapDp), "r"(smem_u32(Gs)), "r"(32), "r"(t0), "r"(b) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (!DZ16 && tile + (int)gridDim.x < n_tiles) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the stores have left the two tiles
        issue_d(tile + gridDim.x);
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// =========================================================================================
// backward 2/3: dx = dx' + dpre[t] . Wcur^T + dpre[t+d] . Wpast^T
// =========================================================================================
struct DxArgs {
  const float* dxn;            // gradient wrt the layer output (nullable for the last layer)
  float* dx;                   // [M][32]
  const unsigned char* img;    // IMG_DX bytes
  int B, T, d;
  int pdl_next;
};

__global__ void __launch_bounds__(256, 2)
block_bwd_dx_umma_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapDn,
                         const __grid_constant__ CUtensorMap mapDx, DxArgs a) {
  // No static shared memory and no alignment slack: with separate in / out tiles two CTAs fill the SM to within 2 KB.
  // The dynamic window then starts at the CTA's (1 KB aligned) shared base; checked, because the swizzle needs it.
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if (smem_u32(smem_raw) & 1023u) __trap();
  unsigned char* P0f = smem;
  unsigned char* P0g = smem + TILE;
  unsigned char* P1f = smem + 2 * TILE;
  unsigned char* P1g = smem + 3 * TILE;
  unsigned char* Sn = smem + 4 * TILE;     // dx' tile (TMA load, read by the threads)
  unsigned char* St = smem + 5 * TILE;     // dx tile on the way out (TMA store).  Separate from Sn: the next dx' load does
                                           // not have to wait until the store has read the tile (a serial ~3000-cycle chain)
  unsigned char* Wb = smem + 6 * TILE;     // Bcf Bcg Bpf Bpg
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wb + IMG_DX);
  uint64_t &bar_tma = bars[0], &bar_n = bars[1], &bar_m1 = bars[2], &bar_w = bars[3];
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;   // thread = (time step, 16-channel half)
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_n, 1);
    mbar_init(&bar_m1, 1);
    mbar_init(&bar_w, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  __syncthreads();
  const int n_tt = (a.T + TM - 1) / TM;
  const int n_tiles = a.B * n_tt;
  const bool have_dxn = a.dxn != nullptr;
  auto issue_loads = [&](int tile) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    mbar_expect_tx(&bar_tma, 4 * TILE);
    tma_load_3d(P0f, &mapP, &bar_tma, 0, t0, b);
    tma_load_3d(P0g, &mapP, &bar_tma, 32, t0, b);
    tma_load_3d(P1f, &mapP, &bar_tma, 0, t0 + a.d, b);     // rows with t + d >= T arrive as zeros
    tma_load_3d(P1g, &mapP, &bar_tma, 32, t0 + a.d, b);
  };
  auto issue_dxn = [&](int tile) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    mbar_expect_tx(&bar_n, TILE);
    tma_load_3d(Sn, &mapDn, &bar_n, 0, t0, b);
  };
  if (tid == 0) {
    mbar_expect_tx(&bar_w, IMG_DX);
    bulk_g2s(Wb, a.img, IMG_DX, &bar_w);
  }
  // dx' (= dx of layer l+1) was written two launches ago: requested before the dependency wait (common.cuh invariant);
  // dpre comes from the direct predecessor
  if (tid == 0 && (int)blockIdx.x < n_tiles && have_dxn) issue_dxn(blockIdx.x);
  pdl_wait();
  if (a.pdl_next) pdl_trigger();
  if (tid == 0 && (int)blockIdx.x < n_tiles) issue_loads(blockIdx.x);
  mbar_wait(&bar_w, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16 * half;
  constexpr uint32_t ID32 = idesc_tf32(128, 32);
  const uint32_t row_off = (uint32_t)r * 128;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    const uint32_t par = it & 1;
    mbar_wait(&bar_tma, par);
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint64_t dp = kmajor_desc(smem_u32(smem + q * TILE)), db = kmajor_desc(smem_u32(Wb + q * 4096));
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem, dp + 2 * k, db + 2 * k, ID32, (q | k) > 0);
      }
      // the previous tile's output store has left the staging tile before anybody (who waits for bar_m1) rewrites it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      mma_commit(&bar_m1);
    }
    float4 xn[4];
    if (have_dxn) {
      mbar_wait(&bar_n, par);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        xn[jj] = *reinterpret_cast<const float4*>(Sn + row_off + ((uint32_t)((4 * half + jj) ^ (r & 7)) << 4));
    } else {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) xn[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    mbar_wait(&bar_m1, par);
    tc_fence_after();
    __syncthreads();      // every thread is past its bar_tma wait before the barrier is re-armed (see bwd_pre)
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) {
      issue_loads(tile + gridDim.x);
      if (have_dxn) issue_dxn(tile + gridDim.x);      // every thread holds its dx' values
    }
    uint32_t ov[16];
    tmem_ld16(lane_addr, ov);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      *reinterpret_cast<float4*>(St + row_off + ((uint32_t)((4 * half + jj) ^ (r & 7)) << 4)) =
          make_float4(xn[jj].x + __uint_as_float(ov[4 * jj]), xn[jj].y + __uint_as_float(ov[4 * jj + 1]),
                      xn[jj].z + __uint_as_float(ov[4 * jj + 2]), xn[jj].w + __uint_as_float(ov[4 * jj + 3]));
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(&mapDx), "r"(smem_u32(St)), "r"(0), "r"(t0), "r"(b) : "memory");   // clipped at the window end
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// =========================================================================================
// backward 3/3: every weight / bias gradient of the layer as ONE GEMM over time
//   D[i][j] = sum_t A[t][i] * B[t][j]          (contraction over time steps)
//   A columns (M = 128): 0-31 x[t] | 32-63 x[t-d] | 64-95 z[t] | 96 ones | 97-127 zero
//   B columns (N = 96) : 0-31 df | 32-63 dg | 64-95 dx'
//   D[0:32 ,0:64] -> filter[1],gate[1]   D[32:64,0:64] -> filter[0],gate[0]   D[64:96,64:96] -> dense
//   D[96,0:64]    -> prebias gradient (per batch element)      D[96,64:96] -> dense_bias
// Both operands are read exactly as the activations lie in HBM ([time][channel], channel contiguous =
// "MN-major"): TMA boxes of [32 steps][32 channels] in the 32-byte-atom swizzle, no transposed copies.
// The dilated past x[t-d] is the same tensor at time coordinate t-d (zero filled for t < d).
// grid = (splits, B): a CTA reduces a contiguous range of 32-step blocks of one batch element.
// =========================================================================================
constexpr int WG_STAGES = 3;
constexpr int WG_ROWS = 64;      // time steps per stage: the kernel is bound by the TMA issue rate of its single producer
                                 // thread (~130 cycles per cp.async.bulk.tensor, measured), so boxes are as tall as smem allows
struct WgArgs {
  float *gwf, *gwg, *gdense, *gprebias, *gdense_bias;
  int B, T, d, is_last, zcol;   // zcol: first column of this layer inside Zcat
  int pdl;                      // launched programmatically inside the single-stream backward chain: wait + trigger
  long long* timeline;          // debug: %globaltimer stamps of CTA 0 (wn_debug_timeline slots 40..47)
};
__device__ __forceinline__ void wg_stamp(const WgArgs& a, int slot) {
  if (a.timeline && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
    a.timeline[40 + slot] = (long long)g;
  }
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(192, 1)
block_wgrad_umma_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ,
                        const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapDn, WgArgs a) {
  constexpr int STG = WG_STAGES;   // (a deeper ring does not help: measured, the kernel is bound by its fixed costs)
  constexpr uint32_t BLK = WG_ROWS * 128;                  // one [64 steps][32 channels] block
  constexpr uint32_t A_BYTES = 4 * BLK, B_BYTES = 3 * BLK, STAGE = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full_bar[STG], empty_bar[STG], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int nkb_total = (a.T + WG_ROWS - 1) / WG_ROWS;
  const int per = (nkb_total + gridDim.x - 1) / gridDim.x;
  const int kb0 = blockIdx.x * per;
  int kb1 = kb0 + per;
  if (kb1 > nkb_total) kb1 = nkb_total;
  const int nk = kb1 - kb0;
  if (nk <= 0) {   // (exiting counts as the trigger; nothing of the predecessor is touched)
    return;
  }
  if (tid == 0) wg_stamp(a, 0);

  // constant block of every stage: column 96 = 1, columns 97..127 = 0; dx' block is zero for the last layer
  for (int i = tid; i < STG * WG_ROWS * 32; i += blockDim.x) {
    const int s = i / (WG_ROWS * 32), rr = (i / 32) % WG_ROWS, cc = i % 32;
    *reinterpret_cast<float*>(smem + s * STAGE + 3 * BLK + swz32(rr, cc)) = (cc == 0) ? 1.0f : 0.0f;
    if (a.is_last) *reinterpret_cast<float*>(smem + s * STAGE + A_BYTES + 2 * BLK + swz32(rr, cc)) = 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < STG; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 128);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) wg_stamp(a, 1);
  if (a.pdl) {
    pdl_wait();
    if (a.pdl == 1) pdl_trigger();      // 2: the successor is a plain launch (end of the chain)
  }
  if (tid == 0) wg_stamp(a, 2);

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = (a.is_last ? 5 : 6) * BLK;      // x, x[t-d], z, [df | dg] (one 4-D box), dx'
      for (int i = 0; i < nk; ++i) {
        const int s = i % STG;
        const uint32_t ph = (i / STG) & 1;
        const int t = (kb0 + i) * WG_ROWS;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], bytes);
        unsigned char* sa = smem + s * STAGE;
        tma_load_3d(sa, &mapX, &full_bar[s], 0, t, b);                      // x[t]
        tma_load_3d(sa + BLK, &mapX, &full_bar[s], 0, t - a.d, b);          // x[t-d]  (zeros for t < d)
        tma_load_3d(sa + 2 * BLK, &mapZ, &full_bar[s], a.zcol, t, b);       // z[t]
        tma_load_4d(sa + A_BYTES, &mapP, &full_bar[s], 0, t, 0, b);         // df | dg as two consecutive blocks
        if (!a.is_last) tma_load_3d(sa + A_BYTES + 2 * BLK, &mapDn, &full_bar[s], 0, t, b);   // dx'
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID = idesc_tf32(128, 96, 1, 1);
      for (int i = 0; i < nk; ++i) {
        const int s = i % STG;
        const uint32_t ph = (i / STG) & 1;
        mbar_wait(&full_bar[s], ph);
        if (i == 0) wg_stamp(a, 3);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE);
        const uint64_t da = mnmajor_desc(sa, BLK), db = mnmajor_desc(sa + A_BYTES, BLK);
#pragma unroll
        for (int k = 0; k < WG_ROWS / 8; ++k) mma_tf32_ss(tmem, da + 64 * k, db + 64 * k, ID, (i | k) > 0);   // +1024 B per K=8
        mma_commit(&empty_bar[s]);
      }
      mma_commit(&done_bar);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(&done_bar, 0);
    if (tid == 64) wg_stamp(a, 4);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < 96; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + c0, v);   // warp-collective: every lane issues it
      float* dst = nullptr;
      if (row < 64) {
        if (c0 < 64) dst = (c0 == 0 ? a.gwf : a.gwg) + ((row < 32 ? 1 : 0) * C + (row & 31)) * C;
      } else if (row < 96) {
        if (c0 == 64 && !a.is_last) dst = a.gdense + (row - 64) * C;
      } else if (row == 96) {
        if (c0 < 64) dst = a.gprebias + (size_t)b * 64 + c0;
        else if (!a.is_last) dst = a.gdense_bias;   // may be null (no biases)
      }
      if (dst) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                     __uint_as_float(v[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) wg_stamp(a, 5);
  if (warp == 1) tmem_dealloc(tmem, 128);
}

// ---- all layers in ONE launch ----
// The per-layer launch above runs 148 CTAs for ~10 stages each and then sends 148 x 12,288 red.adds at the same
// addresses; 50 times per step, inside the dependent backward chain although nothing in the chain reads its output.
// Every layer keeps its own dpre / dx buffers, so once the chain  pre(l) -> dx(l) -> pre(l-1) ...  has finished all
// weight gradients can be computed by one persistent kernel: the (layer, batch element, 64-step block) units are split
// into one contiguous range per SM, the TMA ring runs through the whole range without a bubble, and the accumulators
// are flushed only where (layer, batch element) changes -- ~3 CTAs add to an address instead of 148.
struct WgAllArgs {
  float *gwf, *gwg, *gdense, *gprebias, *gdense_bias;   // bases of the per-layer gradient groups (gdense_bias may be null)
  int L, B, T;
  int last_dense;      // the last layer has a dense output too (stand-alone wn_block_bwd): dx holds L + 1 slots
  int dil[WN_MAX_LAYERS];
};

__global__ void __launch_bounds__(192, 1)
block_wgrad_all_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ,
                       const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapDn,
                       const __grid_constant__ WgAllArgs a) {
  constexpr int STG = WG_STAGES;
  constexpr uint32_t BLK = WG_ROWS * 128;                  // one [64 steps][32 channels] block
  constexpr uint32_t A_BYTES = 4 * BLK, B_BYTES = 3 * BLK, STAGE = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full_bar[STG], empty_bar[STG], done_bar, free_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (a.T + WG_ROWS - 1) / WG_ROWS;
  const long long n_units = (long long)a.L * a.B * nkb;
  const long long per = (n_units + gridDim.x - 1) / gridDim.x;
  const long long u0 = (long long)blockIdx.x * per;
  long long u1 = u0 + per;
  if (u1 > n_units) u1 = n_units;
  if (u0 >= u1) return;

  // constant block of every stage: column 96 = 1, columns 97..127 = 0 (never overwritten by the loads)
  for (int i = tid; i < STG * WG_ROWS * 32; i += blockDim.x) {
    const int s = i / (WG_ROWS * 32), rr = (i / 32) % WG_ROWS, cc = i % 32;
    *reinterpret_cast<float*>(smem + s * STAGE + 3 * BLK + swz32(rr, cc)) = (cc == 0) ? 1.0f : 0.0f;
    *reinterpret_cast<float*>(smem + s * STAGE + A_BYTES + 2 * BLK + swz32(rr, cc)) = 0.f;      // (dx' block: defined before the first load)
  }
  if (tid == 0) {
    for (int s = 0; s < STG; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    mbar_init(&free_bar, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 128);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int per_lb = nkb;      // units per (layer, batch element)

  if (warp == 0) {
    if (lane == 0) {
      uint32_t i = 0;
      for (long long u = u0; u < u1; ++u, ++i) {
        const int lb = (int)(u / per_lb), kb = (int)(u - (long long)lb * per_lb);
        const int l = lb / a.B, b = lb - l * a.B;
        const bool last = (l == a.L - 1) && !a.last_dense;
        const int s = i % STG;
        const uint32_t ph = (i / STG) & 1;
        const int t = kb * WG_ROWS;
        mbar_wait(&empty_bar[s], ph ^ 1);
        // x, x[t-d], z, [df | dg] (one 4-D box), dx' (the last layer has none: its dx' block keeps stale finite data of
        // another layer, which only reaches accumulator columns that are not flushed for the last layer)
        mbar_expect_tx(&full_bar[s], (last ? 5 : 6) * BLK);
        unsigned char* sa = smem + s * STAGE;
        tma_load_3d(sa, &mapX, &full_bar[s], 0, t, lb);                          // x[t] of layer l
        tma_load_3d(sa + BLK, &mapX, &full_bar[s], 0, t - a.dil[l], lb);         // x[t-d]  (zeros for t < d)
        tma_load_3d(sa + 2 * BLK, &mapZ, &full_bar[s], l * C, t, b);             // z[t]
        tma_load_4d(sa + A_BYTES, &mapP, &full_bar[s], 0, t, 0, lb);             // df | dg as two consecutive blocks
        if (!last) tma_load_3d(sa + A_BYTES + 2 * BLK, &mapDn, &full_bar[s], 0, t, lb + a.B);   // dx' = dx of layer l+1
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID = idesc_tf32(128, 96, 1, 1);
      uint32_t i = 0, seg = 0;
      int cur_lb = -1;
      for (long long u = u0; u < u1; ++u, ++i) {
        const int lb = (int)(u / per_lb);
        const bool first = (lb != cur_lb);
        if (first) {
          if (cur_lb >= 0) {
            mma_commit(&done_bar);                       // segment finished: hand the accumulator to the epilogue ...
            mbar_wait(&free_bar, seg & 1);               // ... and wait until it has been read
            tc_fence_after();
            ++seg;
          }
          cur_lb = lb;
        }
        const int s = i % STG;
        const uint32_t ph = (i / STG) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE);
        const uint64_t da = mnmajor_desc(sa, BLK), db = mnmajor_desc(sa + A_BYTES, BLK);
#pragma unroll
        for (int k = 0; k < WG_ROWS / 8; ++k) mma_tf32_ss(tmem, da + 64 * k, db + 64 * k, ID, !(first && k == 0));   // +1024 B per K=8
        mma_commit(&empty_bar[s]);
      }
      mma_commit(&done_bar);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t seg = 0;
    long long u = u0;
    while (u < u1) {
      const int lb = (int)(u / per_lb);
      const int l = lb / a.B, b = lb - l * a.B;
      const bool last = (l == a.L - 1) && !a.last_dense;
      long long ue = (long long)(lb + 1) * per_lb;      // end of this (layer, batch element) inside the range
      if (ue > u1) ue = u1;
      mbar_wait(&done_bar, seg & 1);
      tc_fence_after();
      float* gwf = a.gwf + (size_t)l * 2 * C * C;
      float* gwg = a.gwg + (size_t)l * 2 * C * C;
#pragma unroll 1
      for (int c0 = 0; c0 < 96; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + c0, v);   // warp-collective: every lane issues it
        float* dst = nullptr;
        if (row < 64) {
          if (c0 < 64) dst = (c0 == 0 ? gwf : gwg) + ((row < 32 ? 1 : 0) * C + (row & 31)) * C;
        } else if (row < 96) {
          if (c0 == 64 && !last) dst = a.gdense + (size_t)l * C * C + (row - 64) * C;
        } else if (row == 96) {
          if (c0 < 64) dst = a.gprebias + ((size_t)l * a.B + b) * 64 + c0;
          else if (!last && a.gdense_bias) dst = a.gdense_bias + (size_t)l * C;
        }
        if (dst) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                       __uint_as_float(v[j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&free_bar);
      ++seg;
      u = ue;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

// x: [L][B][T][32] layer inputs; dx: [L][B][T][32] input gradients (dx[l+1] is dx' of layer l); dpre: [L][B][T][64];
// gradient group bases as in the parameter layout (per layer strides 2*C*C, 2*C*C, C*C, B*64, C)
int block_wgrad_all(const float* x, const float* dx, const float* dpre, const float* Zcat, int ldz, float* gwf, float* gwg,
                    float* gdense, float* gprebias, float* gdense_bias, const int* dilations, int L, int B, int T,
                    cudaStream_t st, int last_dense) {
  if (L < 1 || L > WN_MAX_LAYERS) return -1;
  CUtensorMap mX, mZ, mP, mDn;
  int rc = make_map_3d_mn(&mX, x, (int64_t)L * B, T, C, C, WG_ROWS);
  if (rc) return rc;
  rc = make_map_3d_mn(&mZ, Zcat, B, T, ldz, ldz, WG_ROWS);
  if (rc) return rc;
  rc = make_map_4d_mn_blocks(&mP, dpre, (int64_t)L * B, T, 64, WG_ROWS, 2);
  if (rc) return rc;
  rc = make_map_3d_mn(&mDn, dx, (int64_t)(L + (last_dense ? 1 : 0)) * B, T, C, C, WG_ROWS);
  if (rc) return rc;
  WgAllArgs a;
  a.last_dense = last_dense ? 1 : 0;
  a.gwf = gwf; a.gwg = gwg; a.gdense = gdense; a.gprebias = gprebias; a.gdense_bias = gdense_bias;
  a.L = L; a.B = B; a.T = T;
  for (int l = 0; l < WN_MAX_LAYERS; ++l) a.dil[l] = l < L ? dilations[l] : 0;
  const size_t smem = 1024 + WG_STAGES * (7 * WG_ROWS * 128);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(block_wgrad_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const long long n_units = (long long)L * B * ((T + WG_ROWS - 1) / WG_ROWS);
  int grid = sm_count();
  if (grid > n_units) grid = (int)n_units;
  block_wgrad_all_kernel<<<grid, 192, smem, st>>>(mX, mZ, mP, mDn, a);
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_WGRAD);
  return 0;
}

// The three backward kernels of a layer are separate entry points: the weight-gradient GEMM only feeds the
// gradient buffers, so the caller runs it on a side stream next to the dx / next layer's pre kernels.
// fp16 [B][T][ldz] (skip-path gradient of the fp16 chain): box = [TM rows][32 halfs], 64-byte rows, 64B swizzle
static int make_map_dz16(CUtensorMap* m, const void* ptr, int64_t B, int64_t T, int64_t ldz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[3] = {(cuuint64_t)ldz, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)ldz * 2, (cuuint64_t)T * ldz * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)TM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}

// dZcat16 != null: the skip-path gradient is fp16 [M][ldz], scaled by 1 / dz_scale (dZcat is ignored)
int block_bwd_pre_umma(const float* x, const float* dxn, const float* dZcat, const void* dZcat16, float dz_scale, int ldz,
                       int zcol, float* dpre, const unsigned char* img_pre, const float* prebias, int B, int T, int d,
                       int is_last, int pdl_next, cudaStream_t st) {
  const int n_tiles = B * ((T + TM - 1) / TM);
  int grid = n_tiles;
  const int cap = 2 * sm_count();
  if (grid > cap) grid = cap;
  CUtensorMap mX, mDn, mDz;
  int rc = make_map_3d(&mX, x, B, T, C, C, TM);
  if (rc) return rc;
  rc = make_map_3d(&mDn, is_last ? x : dxn, B, T, C, C, TM);
  if (rc) return rc;
  rc = dZcat16 ? make_map_dz16(&mDz, dZcat16, B, T, ldz) : make_map_3d(&mDz, dZcat, B, T, ldz, ldz, TM);
  if (rc) return rc;
  CUtensorMap mDp;
  rc = make_map_3d(&mDp, dpre, B, T, 64, 64, TM);
  if (rc) return rc;
  PreArgs a;
  a.dpre = dpre; a.img = img_pre; a.prebias = prebias; a.B = B; a.T = T; a.d = d;
  a.is_last = is_last; a.zcol = zcol; a.pdl_next = pdl_next; a.dz_scale = dz_scale;
  const size_t smem = 1024 + 4 * TILE + IMG_PRE + (dZcat16 ? TILE + TM * 64 : 0);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(block_bwd_pre_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + 4 * TILE + IMG_PRE));
    cudaFuncSetAttribute(block_bwd_pre_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + 5 * TILE + IMG_PRE + TM * 64));
    attr = true;
  }
  auto kernel = dZcat16 ? block_bwd_pre_umma_kernel<true> : block_bwd_pre_umma_kernel<false>;
  // Plain stream-ordered launches in the backward chain: with programmatic dependent launch next to the cross-stream
  // events of the side-stream weight-gradient kernels, consumers were observed to start on half-written gradients
  // (tools/sweep_impls.py, tools/debug_case.py); the forward chain (no events) keeps PDL.
  if (pdl_next >= 0) {      // whole backward chain on one stream, no events in between: PDL as in the forward chain
    a.pdl_next = pdl_next;
    cudaError_t e = launch_pdl(kernel, dim3(grid), dim3(256), smem, st, mX, mDn, mDz, mDp, a);
    if (e != cudaSuccess) return (int)e;
  } else {
    a.pdl_next = 0;
    kernel<<<grid, 256, smem, st>>>(mX, mDn, mDz, mDp, a);
    WN_CHECK_LAUNCH();
  }
  prof_mark(st, PT_BLOCK_BWD_PRE);
  return 0;
}

int block_wgrad_umma(const float* x, const float* dxn, const float* dpre, const float* Zcat, int ldz, int zcol,
                     float* gwf, float* gwg, float* gdense, float* gprebias, float* gdense_bias, int B, int T, int d,
                     int is_last, int pdl, cudaStream_t st) {
  CUtensorMap mX, mZ, mP, mDn;
  int rc = make_map_3d_mn(&mX, x, B, T, C, C, WG_ROWS);
  if (rc) return rc;
  rc = make_map_3d_mn(&mZ, Zcat, B, T, ldz, ldz, WG_ROWS);
  if (rc) return rc;
  rc = make_map_4d_mn_blocks(&mP, dpre, B, T, 64, WG_ROWS, 2);
  if (rc) return rc;
  rc = make_map_3d_mn(&mDn, is_last ? x : dxn, B, T, C, C, WG_ROWS);
  if (rc) return rc;
  WgArgs a;
  a.gwf = gwf; a.gwg = gwg; a.gdense = gdense; a.gprebias = gprebias; a.gdense_bias = gdense_bias; a.B = B; a.T = T;
  a.d = d; a.is_last = is_last; a.zcol = zcol; a.pdl = pdl; a.timeline = g_timeline;
  const size_t smem = 1024 + WG_STAGES * (7 * WG_ROWS * 128);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(block_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const int nkb = (T + WG_ROWS - 1) / WG_ROWS;
  int splits = sm_count() / (B > 0 ? B : 1);
  if (splits < 1) splits = 1;
  if (splits > nkb) splits = nkb;
  if (pdl) {
    cudaError_t e = launch_pdl(block_wgrad_umma_kernel, dim3(splits, B), dim3(192), smem, st, mX, mZ, mP, mDn, a);
    if (e != cudaSuccess) return (int)e;
  } else {   // plain (fully stream-ordered) launch: on the side stream this kernel follows a cross-stream event wait
    block_wgrad_umma_kernel<<<dim3(splits, B), 192, smem, st>>>(mX, mZ, mP, mDn, a);
    WN_CHECK_LAUNCH();
  }
  prof_mark(st, PT_BLOCK_WGRAD);
  return 0;
}

int block_bwd_dx_umma(const float* dxn, const float* dpre, float* dx, const unsigned char* img_dx, int B, int T, int d,
                      int is_last, int pdl_next, cudaStream_t st) {
  const int n_tiles = B * ((T + TM - 1) / TM);
  int grid = n_tiles;
  const int cap = 2 * sm_count();
  if (grid > cap) grid = cap;
  CUtensorMap mP, mDn, mDx;
  int rc = make_map_3d(&mP, dpre, B, T, 64, 64, TM);
  if (rc) return rc;
  rc = make_map_3d(&mDx, dx, B, T, C, C, TM);
  if (rc) return rc;
  mDn = mDx;
  if (!is_last) {
    rc = make_map_3d(&mDn, dxn, B, T, C, C, TM);
    if (rc) return rc;
  }
  DxArgs a;
  a.dxn = is_last ? nullptr : dxn; a.dx = dx; a.img = img_dx; a.B = B; a.T = T; a.d = d; a.pdl_next = pdl_next;
  const size_t smem = 6 * TILE + IMG_DX + 64;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(block_bwd_dx_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  if (pdl_next >= 0) {
    a.pdl_next = pdl_next;
    cudaError_t e = launch_pdl(block_bwd_dx_umma_kernel, dim3(grid), dim3(256), smem, st, mP, mDn, mDx, a);
    if (e != cudaSuccess) return (int)e;
  } else {
    a.pdl_next = 0;
    block_bwd_dx_umma_kernel<<<grid, 256, smem, st>>>(mP, mDn, mDx, a);
    WN_CHECK_LAUNCH();
  }
  prof_mark(st, PT_BLOCK_BWD_DX);
  return 0;
}

int block_umma_set_trap_info(unsigned int* p) { return umma::set_trap_info_tu(p); }

}  // namespace wn
