// Backward of the gated residual blocks, second generation: ONE persistent, flag-ordered kernel for the whole stack
// (pre-activation gradient and input gradient of a tile fused), fp16 split-row operands on tcgen05 kind::f16, plus the
// weight gradients of all layers as one GEMM over time on the same fp16 tiles.
// Reference: TF autodiff of wavenet/model.py:236-330 (_create_dilation_layer) and wavenet/ops.py:46-62 (causal_conv).
//
// Per layer l (input x_l, output gradient dx' = dx_{l+1}, skip-path gradient dz_skip from the post-processing GEMMs):
//     pre   = x[t-d].W0 + x[t].W1 + prebias            (recomputed: the same six-MMA split product as the forward)
//     dz    = dz_skip + dx'.Wd^T
//     dpre  = [ dz.sig(g).(1 - tanh(f)^2) | dz.tanh(f).sig(g).(1 - sig(g)) ]
//     dx[t] = dx'[t] + dpre[t].W1^T + dpre[t+d].W0^T
// Data layout: every activation-sized tensor of the backward pass is fp16 in a power-of-two scaled domain (the gradient
// chain of the post-processing GEMMs already is): dpre as [df 32 | dg 32] rows of 64 halfs, dx as split rows
// [hi 32 | lo 32] (22 significant bits, like the forward residual stream); x comes from the split rows the forward
// chain kept for every layer.  A row of 64 halfs is one 128-byte swizzle row, so the same shared-memory tile is
//   * a K-major operand (contraction over channels: the products above), and
//   * an MN-major operand (contraction over time: the weight gradients)        -- no transposed or re-swizzled copies.
// Work item = (layer, batch element, 128-step tile), claimed from one global counter in (layer descending, time
// descending) order; tile (l, t0) needs dx' of (l+1, t0) and dpre of the same layer at [t0+d, t0+d+128), i.e. tiles
// later in time = earlier in the order, so the kernel is deadlock-free for any number of resident CTAs.
//   warps 0-7  epilogue (thread = one time step x 16 channels)
//   warp  8    issuer: claims work, polls flags, TMA loads, tcgen05 MMAs, weight images
//   warp  9    publisher: TMA stores of dpre and dx, then (stores complete) the tile's flags
// The shifted operand dpre[t+d]: the tile's own dpre rows sit in shared memory (P), the rows of the tile(s) after it are
// loaded right behind them (Pn), and the MMA reads [P | Pn] from row min(d, 128) on: a descriptor start address in the
// middle of the swizzle pattern (matrix base offset = row & 7).  WN_BWD_SHIFT=global takes the own rows through L2 instead.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>

#include "../../include/wavenet_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "umma_common.cuh"
#include "split_common.cuh"

namespace wn {
using namespace umma;

namespace {
constexpr uint32_t IMG_B = 3 * 4096;          // WdT [32 c][hi r | lo r] | Bcur [32 r][df c | dg c] | Bpast
constexpr uint32_t W_BYTES = 16384 + IMG_B;   // W0cat | W1cat (forward image) | backward image
constexpr int BC_THREADS = 320;

__device__ __forceinline__ __half sat_half(float x) { return __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f)); }
__device__ __forceinline__ void split_sat(float x, __half& h, __half& l) {
  x = fminf(fmaxf(x, -65504.f), 65504.f);
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}
// fp16 [B][T][cols]: box = [rows][64 halfs] (128-byte rows, 128B swizzle) at (col, t, b)
static int make_map_h64(CUtensorMap* m, const void* ptr, int64_t B, int64_t T, int64_t cols, int64_t ld, int rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}
}  // namespace

// ---- backward weight images: per layer  WdT | Bcur | Bpast  (fp16, swizzled K-major rows of 64 halfs) ----
__global__ void block_bwd_h_images_kernel(unsigned char* __restrict__ img, const float* __restrict__ filter,
                                          const float* __restrict__ gate, const float* __restrict__ dense) {
  const int l = blockIdx.x;
  unsigned char* base = img + (size_t)l * IMG_B;
  const float* wf = filter + (size_t)l * 2 * C * C;
  const float* wg = gate + (size_t)l * 2 * C * C;
  const float* wd = dense + (size_t)l * C * C;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {      // WdT: n = dilation channel c, k = residual channel r
    const int n = i / 32, k = i % 32;
    __half h, lo;
    split_h(wd[n * C + k], h, lo);
    *reinterpret_cast<__half*>(base + swzh(n, k)) = h;
    *reinterpret_cast<__half*>(base + swzh(n, 32 + k)) = lo;
  }
  for (int i = threadIdx.x; i < 2 * 32 * 64; i += blockDim.x) {  // B[tap]: n = residual channel r, k = [df c | dg c]
    const int which = i / (32 * 64), n = (i / 64) % 32, k = i % 64;
    const int tap = which == 0 ? 1 : 0;      // Bcur multiplies dpre[t] (tap 1 = current sample), Bpast dpre[t+d] (tap 0)
    const float w = k < 32 ? wf[(tap * C + n) * C + k] : wg[(tap * C + n) * C + (k - 32)];
    *reinterpret_cast<__half*>(base + 4096 + which * 4096 + swzh(n, k)) = __float2half_rn(w);
  }
}
int64_t block_bwd_h_images_bytes(int L) { return (int64_t)L * IMG_B; }
int block_bwd_h_images(unsigned char* img, const float* filter, const float* gate, const float* dense, int L, cudaStream_t st) {
  block_bwd_h_images_kernel<<<L, 256, 0, st>>>(img, filter, gate, dense);
  WN_CHECK_LAUNCH();
  return 0;
}

// split rows [M][hi 32 | lo 32] -> fp32 [M][32] * scale   (the gradient wrt the first layer's input leaves the fp16 domain)
__global__ void unsplit_rows_kernel(const __half* __restrict__ xs, float* __restrict__ x, int64_t n4, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int64_t m = i >> 3;
    const int c4 = (int)(i & 7) * 4;
    const __half* row = xs + m * 64;
    const uint2 hv = *reinterpret_cast<const uint2*>(row + c4), lv = *reinterpret_cast<const uint2*>(row + 32 + c4);
    const __half2* h = reinterpret_cast<const __half2*>(&hv);
    const __half2* l = reinterpret_cast<const __half2*>(&lv);
    const float2 h0 = __half22float2(h[0]), h1 = __half22float2(h[1]), l0 = __half22float2(l[0]), l1 = __half22float2(l[1]);
    reinterpret_cast<float4*>(x)[i] = make_float4((h0.x + l0.x) * scale, (h0.y + l0.y) * scale, (h1.x + l1.x) * scale, (h1.y + l1.y) * scale);
  }
}
int unsplit_rows(const void* xs, float* x, int64_t M, float scale, cudaStream_t st) {
  const int64_t n4 = M * 8;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  unsplit_rows_kernel<<<(int)blocks, 256, 0, st>>>((const __half*)xs, x, n4, scale);
  WN_CHECK_LAUNCH();
  return 0;
}

// =====================================================================================================================
// the chain kernel
// =====================================================================================================================
struct BwdChainArgs {
  const unsigned char* img_f;      // [L] forward weight images (IMG_H bytes each; W0cat | W1cat are used)
  const unsigned char* img_b;      // [L] backward weight images (IMG_B bytes each)
  const float* prebias;            // [L][B][64]
  unsigned int* flags;             // [L][n_tiles] dpre published | [L][n_tiles] dx published | work counter; zeroed before the launch
  int L, B, T, n_tt;
  int last_dense;                  // the last layer has an output gradient too (stand-alone wn_block_bwd)
  int base_off;                    // (probe) set the descriptor's matrix base offset field to the start row & 7
  int shift_global;                // dilations that are not multiples of 8 (and < 128): own dpre rows through L2
  float cs;                        // the skip-path gradient is multiplied by cs on its way into the chain's scaled domain
  long long* timeline;             // debug: cycles CTA 0 spent in each kind of wait
  int dil[WN_MAX_LAYERS];
};
static long long* g_timeline_b = nullptr;
void set_bwd_h_timeline(long long* p) { g_timeline_b = p; }

__global__ void __launch_bounds__(BC_THREADS, 2)
block_bwd_chain_kernel(const __grid_constant__ CUtensorMap mapXS, const __grid_constant__ CUtensorMap mapDX,
                       const __grid_constant__ CUtensorMap mapDz, const __grid_constant__ CUtensorMap mapP,
                       const __grid_constant__ BwdChainArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Xc = smem;                 // split rows x[t]
  unsigned char* P = smem + TILE;           // dpre of this tile [128][df 32 | dg 32] fp16
  unsigned char* Xp = smem + 2 * TILE;      // split rows x[t-d]; later Pn: the dpre rows that follow this tile's ([P | Pn] is contiguous)
  unsigned char* Dn = smem + 3 * TILE;      // split rows dx' (A operand of dx'.Wd^T, residual term); later the dx rows on their way out
  unsigned char* Dz = smem + 4 * TILE;      // skip-path gradient [128][32 halfs], 64B swizzle
  unsigned char* W0 = Dz + TM * 64;         // [64][hi|lo] past tap (filter | gate)      } forward image
  unsigned char* W1 = W0 + 8192;            // current tap                               }
  unsigned char* WdT = W1 + 8192;           // [32 c][hi r | lo r]                        } backward image
  unsigned char* Bc = WdT + 4096;           // [32 r][df c | dg c], tap 1
  unsigned char* Bp = Bc + 4096;            // tap 0
  __shared__ __align__(8) uint64_t bar_x, bar_d, bar_m1, bar_p, bar_pn, bar_m2, bar_o, bar_pfree, bar_sfree, bar_w;
  __shared__ uint32_t tmem_slot;
  __shared__ int item_s[4];                 // work item of tile i in item_s[i & 3] (-1: no more work), published through bar_x
  __shared__ float pb_s[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_x, 1);
    mbar_init(&bar_d, 1);
    mbar_init(&bar_m1, 1);
    mbar_init(&bar_p, 256);
    mbar_init(&bar_pn, 1);
    mbar_init(&bar_m2, 1);
    mbar_init(&bar_o, 256);
    mbar_init(&bar_pfree, 1);
    mbar_init(&bar_sfree, 1);
    mbar_init(&bar_w, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;          // columns 0-63 pre-activations (f | g), 64-95 dx'.Wd^T, 96-127 the dx products
  const int n_tiles = a.B * a.n_tt;
  const int n_items = a.L * n_tiles;
  unsigned int* flagP = a.flags;
  unsigned int* flagX = a.flags + (size_t)n_items;
  // item -> (layer, batch element, tile): layers descending, time descending
  auto decode = [&](int item, int& l, int& b, int& tt) {
    const int q = item / n_tiles, j = item - q * n_tiles;
    l = a.L - 1 - q;
    b = j / a.n_tt;
    tt = a.n_tt - 1 - (j - b * a.n_tt);
  };

  if (warp == 8) {
    // ------------------------------------------------------------------------------------------------ issuer
    if (lane == 0) {
      constexpr uint32_t ID64 = idesc_f16(128, 64), ID32 = idesc_f16(128, 32);
      const uint64_t dXc = kmajor_desc(smem_u32(Xc)), dXp = kmajor_desc(smem_u32(Xp)), dDn = kmajor_desc(smem_u32(Dn));
      const uint64_t dP = kmajor_desc(smem_u32(P));
      const uint64_t dW0 = kmajor_desc(smem_u32(W0)), dW1 = kmajor_desc(smem_u32(W1)), dWdT = kmajor_desc(smem_u32(WdT));
      const uint64_t dBc = kmajor_desc(smem_u32(Bc)), dBp = kmajor_desc(smem_u32(Bp));
      unsigned int* counter = a.flags + 2 * (size_t)n_items;
      auto claim = [&]() -> int {
        const unsigned int w = atomicAdd(counter, 1u);
        return w < (unsigned int)n_items ? (int)w : -1;
      };
      auto load_weights = [&](int l) {
        mbar_expect_tx(&bar_w, W_BYTES);
        bulk_g2s(W0, a.img_f + (size_t)l * IMG_H, 16384, &bar_w);
        bulk_g2s(WdT, a.img_b + (size_t)l * IMG_B, IMG_B, &bar_w);
      };
      auto has_dn_of = [&](int l) { return l < a.L - 1 || a.last_dense; };
      // x tiles: written by the forward pass, always ready
      auto load_x = [&](int l, int b, int t0) {
        mbar_expect_tx(&bar_x, 2 * TILE);      // (release: publishes item_s to the waiters of this phase)
        tma_load_3d(Xc, &mapXS, &bar_x, 0, t0, l * a.B + b);
        tma_load_3d(Xp, &mapXS, &bar_x, 0, t0 - a.dil[l], l * a.B + b);      // rows before the window start arrive as zeros
      };
      // skip-path gradient (ready) and dx' of the layer above (needs that tile's flag)
      auto load_d = [&](int l, int b, int tt) {
        const bool hd = has_dn_of(l);
        if (hd && l < a.L - 1) wait_flag(flagX + (size_t)(l + 1) * n_tiles + (size_t)b * a.n_tt + tt);      // (last_dense: dx' of the last layer is an input)
        mbar_expect_tx(&bar_d, TM * 64 + (hd ? TILE : 0));
        tma_load_3d(Dz, &mapDz, &bar_d, l * C, tt * TM, b);
        if (hd) tma_load_3d(Dn, &mapDX, &bar_d, 0, tt * TM, (l + 1) * a.B + b);
      };
      long long c_x = 0, c_d = 0, c_fl = 0, c_p = 0, c_pn = 0, c_m2 = 0, c_sf = 0, c_w = 0, t_a, t_start = clock64();
      int item = claim();
      item_s[0] = item;
      uint32_t i = 0, wphase = 0;
      int l = 0, b = 0, tt = 0;
      if (item < 0) {
        mbar_arrive(&bar_x);
      } else {
        decode(item, l, b, tt);
        load_weights(l);
        load_x(l, b, tt * TM);
        load_d(l, b, tt);
        mbar_wait(&bar_w, 0);
      }
      while (item >= 0) {
        const uint32_t par = i & 1;
        const int t0 = tt * TM, d = a.dil[l];
        const bool hd = has_dn_of(l);
        t_a = clock64();
        mbar_wait(&bar_x, par);
        c_x += clock64() - t_a;
        tc_fence_after();
        mma_split(tmem, dXp, dW0, ID64, true);      // x[t-d] . W[0]
        mma_split(tmem, dXc, dW1, ID64, false);     // x[t]   . W[1]
        t_a = clock64();
        mbar_wait(&bar_d, par);
        c_d += clock64() - t_a;
        tc_fence_after();
        if (hd) mma_split(tmem + 64, dDn, dWdT, ID32, true);      // dx' . Wd^T
        mma_commit(&bar_m1);
        // the dpre rows behind this tile: [t0 + 128, ...) for d < 128 (the own rows are in P), [t0 + d, ...) otherwise
        const bool via_l2 = a.shift_global && d < TM && (d & 7);
        const int tpn = via_l2 ? t0 + d : t0 + (d > TM ? d : TM);
        const int row0 = via_l2 ? TM : (d < TM ? d : TM);      // first row of [P | Pn] the shifted operand reads
        mbar_wait(&bar_m1, par);      // the x tiles have been read: Xp becomes Pn
        t_a = clock64();
        {
          const unsigned int* f = flagP + (size_t)l * n_tiles + (size_t)b * a.n_tt;
          const int ta = tpn / TM, tb = (tpn + TM - 1) / TM;
          if (ta < a.n_tt) wait_flag(f + ta);
          if (tb < a.n_tt && tb != ta) wait_flag(f + tb);
        }
        c_fl += clock64() - t_a;
        mbar_expect_tx(&bar_pn, TILE);
        tma_load_3d(Xp, &mapP, &bar_pn, 0, tpn, l * a.B + b);      // rows at or past the window end arrive as zeros
        t_a = clock64();
        mbar_wait(&bar_p, par);       // P staged by the epilogue threads
        c_p += clock64() - t_a;
        t_a = clock64();
        mbar_wait(&bar_pn, par);
        c_pn += clock64() - t_a;
        tc_fence_after();
        {
          const uint32_t sh = smem_u32(P) + (uint32_t)row0 * 128u;
          const uint64_t dSh = kmajor_desc(sh) | (a.base_off ? ((uint64_t)((sh >> 7) & 7u) << 49) : 0ull);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16_ss(tmem + 96, dP + 2 * k, dBc + 2 * k, ID32, k > 0);      // dpre[t]   . W[1]^T
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16_ss(tmem + 96, dSh + 2 * k, dBp + 2 * k, ID32, 1u);         // dpre[t+d] . W[0]^T
        }
        mma_commit(&bar_m2);
        // ---- next item ----
        const int nx = claim();
        item_s[(i + 1) & 3] = nx;
        t_a = clock64();
        mbar_wait(&bar_m2, par);      // P / Pn / Bc / Bp have been read
        c_m2 += clock64() - t_a;
        if (nx < 0) {
          mbar_arrive(&bar_x);        // completes the next phase without data: the other warps see "no more work"
          break;
        }
        int nl, nb, ntt;
        decode(nx, nl, nb, ntt);
        if (nl != l) {
          load_weights(nl);
          wphase ^= 1;
        }
        load_x(nl, nb, ntt * TM);
        t_a = clock64();
        mbar_wait(&bar_sfree, par);   // the dx store has read Dn; every thread has read Dz
        c_sf += clock64() - t_a;
        load_d(nl, nb, ntt);
        if (nl != l) {
          t_a = clock64();
          mbar_wait(&bar_w, wphase);
          c_w += clock64() - t_a;
        }
        item = nx; l = nl; b = nb; tt = ntt;
        ++i;
      }
      if (a.timeline && blockIdx.x == 0) {
        a.timeline[0] = clock64() - t_start; a.timeline[1] = c_x; a.timeline[2] = c_d; a.timeline[3] = c_fl; a.timeline[4] = c_p;
        a.timeline[5] = c_pn; a.timeline[6] = c_m2; a.timeline[7] = c_sf; a.timeline[8] = c_w; a.timeline[9] = i; a.timeline[10] = gridDim.x;
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------------------------------------ publisher
    if (lane == 0) {
      for (uint32_t i = 0;; ++i) {
        const uint32_t par = i & 1;
        mbar_wait(&bar_p, par);
        const int item = item_s[i & 3];
        if (item < 0) break;
        int l, b, tt;
        decode(item, l, b, tt);
        const size_t fidx = (size_t)l * n_tiles + (size_t)b * a.n_tt + tt;
        tma_store_3d(&mapP, P, 0, tt * TM, l * a.B + b);      // rows past the end of the window are clipped by the tensor map
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(&bar_pfree);
        bulk_wait0();                 // the store has completed and is visible to this thread ...
        __threadfence();
        st_release(flagP + fidx, 1u);      // ... publish the tile's dpre
        mbar_wait(&bar_o, par);
        tma_store_3d(&mapDX, Dn, 0, tt * TM, l * a.B + b);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(&bar_sfree);
        bulk_wait0();
        __threadfence();
        st_release(flagX + fidx, 1u);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------------ epilogue
    const int r = tid & 127, half = tid >> 7;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16 * half;
    const uint32_t row_off = (uint32_t)r * 128;
    const uint32_t ch0 = (uint32_t)((2 * half) ^ (r & 7)) << 4, ch1 = (uint32_t)((2 * half + 1) ^ (r & 7)) << 4;
    const uint32_t cl0 = (uint32_t)((4 + 2 * half) ^ (r & 7)) << 4, cl1 = (uint32_t)((5 + 2 * half) ^ (r & 7)) << 4;
    int key = -1;
    for (uint32_t i = 0;; ++i) {
      const uint32_t par = i & 1;
      mbar_wait(&bar_x, par);
      const int item = item_s[i & 3];
      if (item < 0) {
        // hand the end marker on to the publisher once it has consumed the previous phase of bar_p (its bar_pfree
        // arrival follows its bar_p wait): a parity wait that falls two phases behind never returns
        if (i > 0) mbar_wait(&bar_pfree, (i - 1) & 1);
        mbar_arrive(&bar_p);
        break;
      }
      int l, b, tt;
      decode(item, l, b, tt);
      const bool hd = l < a.L - 1 || a.last_dense;
      if (l * a.B + b != key) {      // uniform over the epilogue warps: bias / conditioning row of this (layer, batch element)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < 64) pb_s[tid] = a.prebias[((size_t)l * a.B + b) * 64 + tid];
        key = l * a.B + b;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(&bar_d, par);
      float dz[16];
      {
        const unsigned char* zr = Dz + (uint32_t)r * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(zr + ((uint32_t)((2 * half + c) ^ ((r >> 1) & 3)) << 4));
          const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(h[q]);
            dz[8 * c + 2 * q] = f.x * a.cs; dz[8 * c + 2 * q + 1] = f.y * a.cs;
          }
        }
      }
      mbar_wait(&bar_m1, par);
      tc_fence_after();
      if (hd) {
        uint32_t av[16];
        tmem_ld16(lane_addr + 64, av);
#pragma unroll
        for (int q = 0; q < 16; ++q) dz[q] += __uint_as_float(av[q]);
      }
      const bool valid = (tt * TM + r) < a.T;
      __align__(16) __half dfh[16], dgh[16];
      {
        uint32_t fv[16], gv[16];
        tmem_ld16(lane_addr + 0, fv);
        tmem_ld16(lane_addr + 32, gv);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float tf, sg;
          gated_parts_fast(__uint_as_float(fv[q]) + pb_s[16 * half + q], __uint_as_float(gv[q]) + pb_s[32 + 16 * half + q], tf, sg);
          const float dzv = valid ? dz[q] : 0.f;      // rows past the window end act as the (zero) future of the rows before them
          dfh[q] = sat_half(dzv * sg * (1.f - tf * tf));
          dgh[q] = sat_half(dzv * tf * sg * (1.f - sg));
        }
      }
      if (i > 0) mbar_wait(&bar_pfree, (i - 1) & 1);      // the previous tile's dpre store has left P
      *reinterpret_cast<uint4*>(P + row_off + ch0) = *reinterpret_cast<const uint4*>(dfh);
      *reinterpret_cast<uint4*>(P + row_off + ch1) = *reinterpret_cast<const uint4*>(dfh + 8);
      *reinterpret_cast<uint4*>(P + row_off + cl0) = *reinterpret_cast<const uint4*>(dgh);
      *reinterpret_cast<uint4*>(P + row_off + cl1) = *reinterpret_cast<const uint4*>(dgh + 8);
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(&bar_p);
      // ---- dx = dx' + the two products ----
      float dx[16];
      if (hd) {
        const uint4 h0 = *reinterpret_cast<const uint4*>(Dn + row_off + ch0), h1 = *reinterpret_cast<const uint4*>(Dn + row_off + ch1);
        const uint4 l0 = *reinterpret_cast<const uint4*>(Dn + row_off + cl0), l1 = *reinterpret_cast<const uint4*>(Dn + row_off + cl1);
        const __half2* hh0 = reinterpret_cast<const __half2*>(&h0);
        const __half2* hh1 = reinterpret_cast<const __half2*>(&h1);
        const __half2* ll0 = reinterpret_cast<const __half2*>(&l0);
        const __half2* ll1 = reinterpret_cast<const __half2*>(&l1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a0 = __half22float2(hh0[q]), b0 = __half22float2(ll0[q]);
          const float2 a1 = __half22float2(hh1[q]), b1 = __half22float2(ll1[q]);
          dx[2 * q] = a0.x + b0.x; dx[2 * q + 1] = a0.y + b0.y;
          dx[8 + 2 * q] = a1.x + b1.x; dx[8 + 2 * q + 1] = a1.y + b1.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) dx[q] = 0.f;
      }
      mbar_wait(&bar_m2, par);
      tc_fence_after();
      {
        uint32_t ov[16];
        tmem_ld16(lane_addr + 96, ov);
#pragma unroll
        for (int q = 0; q < 16; ++q) dx[q] += __uint_as_float(ov[q]);
      }
      __align__(16) __half xh[16], xl[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) split_sat(dx[q], xh[q], xl[q]);
      *reinterpret_cast<uint4*>(Dn + row_off + ch0) = *reinterpret_cast<const uint4*>(xh);
      *reinterpret_cast<uint4*>(Dn + row_off + ch1) = *reinterpret_cast<const uint4*>(xh + 8);
      *reinterpret_cast<uint4*>(Dn + row_off + cl0) = *reinterpret_cast<const uint4*>(xl);
      *reinterpret_cast<uint4*>(Dn + row_off + cl1) = *reinterpret_cast<const uint4*>(xl + 8);
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(&bar_o);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 128);
}

static int bwd_shift_mode() {      // WN_BWD_SHIFT=global: dilations that are not multiples of 8 take the own dpre rows through L2; =baseoff: probe
  static int v = -1;
  if (v < 0) { const char* e = getenv("WN_BWD_SHIFT"); v = (e && strcmp(e, "global") == 0) ? 1 : (e && strcmp(e, "baseoff") == 0) ? 2 : 0; }
  return v;
}

// xs: [L][B][T][hi 32 | lo 32] layer inputs (the forward chain's ring with L slots); dxs: [L (+1)][B][T][hi | lo] input
// gradients in the scaled domain (slot l+1 is dx' of layer l; slot L exists only with last_dense); p16: [L][B][T][64]
// dpre; dz16: [B][T][ldz] skip-path gradient; flags: 2 * L * B * ceil(T/128) + 1 words.
int block_bwd_chain(const void* xs, void* dxs, void* p16, const void* dz16, int ldz, float cs, const unsigned char* img_f,
                    const unsigned char* img_b, const float* prebias, const int* dilations, int L, int B, int T,
                    unsigned int* flags, cudaStream_t st, int last_dense) {
  if (L < 1 || L > WN_MAX_LAYERS) return -1;
  CUtensorMap mapXS, mapDX, mapDz, mapP;
  int rc = make_map_split(&mapXS, (const __half*)xs, (int64_t)L * B, T);
  if (rc) return rc;
  rc = make_map_split(&mapDX, (const __half*)dxs, (int64_t)(L + (last_dense ? 1 : 0)) * B, T);
  if (rc) return rc;
  rc = make_map_z16(&mapDz, (const __half*)dz16, B, T, ldz);
  if (rc) return rc;
  rc = make_map_split(&mapP, (const __half*)p16, (int64_t)L * B, T);
  if (rc) return rc;
  BwdChainArgs a;
  a.img_f = img_f; a.img_b = img_b; a.prebias = prebias; a.flags = flags;
  a.L = L; a.B = B; a.T = T; a.n_tt = (T + TM - 1) / TM;
  a.last_dense = last_dense ? 1 : 0; a.shift_global = bwd_shift_mode() == 1 ? 1 : 0; a.base_off = bwd_shift_mode() == 2 ? 1 : 0; a.cs = cs;
  a.timeline = g_timeline_b;
  for (int l = 0; l < WN_MAX_LAYERS; ++l) a.dil[l] = l < L ? dilations[l] : 0;
  const size_t smem = 1024 + 4 * TILE + TM * 64 + W_BYTES;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(block_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -5;
    cudaFuncSetAttribute(block_bwd_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  const int64_t n_items = (int64_t)L * B * a.n_tt;
  if (n_items >= (1ll << 30)) return -1;
  int grid = 2 * sm_count();
  if (grid > B * a.n_tt) grid = B * a.n_tt;
  cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)(2 * n_items + 1) * sizeof(unsigned int), st);
  if (e != cudaSuccess) return (int)e;
  block_bwd_chain_kernel<<<grid, BC_THREADS, smem, st>>>(mapXS, mapDX, mapDz, mapP, a);
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_BWD_PRE);
  return 0;
}
int64_t block_bwd_chain_flag_words(int L, int B, int T) { return 2 * (int64_t)L * B * ((T + TM - 1) / TM) + 1; }

// =====================================================================================================================
// weight gradients of ALL layers: one GEMM over time per (layer, batch element) on the fp16 tiles
//   D1[i][j] = sum_t [x | x[t-d]][t][i] * dpre[t][j]            i: x hi 0-31, x lo 32-63, x[t-d] hi 64-95, lo 96-127
//   D2[i][j] = sum_t [z_l | z_l+1 | 1 | 0..][t][i] * [dpre | dx'][t][j]     rows 0-31: z.dx' -> dense; row 64: bias sums
// Operands are MN-major blocks of [64 time steps][64 halfs] exactly as they lie in memory (plain 128B swizzle, LBO =
// distance between blocks, SBO 1024, 2048 B per K = 16 step).  The hi and lo rows of x add into the same gradient
// element (red.add), so x enters with its full split precision; z comes from the fp16 Zcat the skip GEMM reads (a box of
// 64 columns starting at the layer's column: the upper half belongs to the next layer and its products are not flushed).
// The (layer, batch element, 64-step block) units are cut into one contiguous range per SM as in block_wgrad_all.
// =====================================================================================================================
constexpr int WGH_STAGES = 4;
constexpr int WGH_ROWS = 64;
constexpr uint32_t WGH_BLK = WGH_ROWS * 128;      // one [64 steps][64 halfs] block
constexpr uint32_t WGH_STAGE = 5 * WGH_BLK;       // x | x[t-d] | z | dpre | dx'

struct WgHArgs {
  float *gwf, *gwg, *gdense, *gprebias, *gdense_bias;   // bases of the per-layer gradient groups (gdense_bias may be null)
  int L, B, T;
  int last_dense;
  float scale;                 // out of the chain's scaled domain
  int dil[WN_MAX_LAYERS];
};

__device__ __forceinline__ void red_add_v4h(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint64_t mn16_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1)
block_wgrad_h_all_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ,
                         const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapDn,
                         const __grid_constant__ WgHArgs a) {
  constexpr int STG = WGH_STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Ones = smem + STG * WGH_STAGE;      // constant block: column 0 = 1, the rest 0
  __shared__ __align__(8) uint64_t full_bar[STG], empty_bar[STG], done_bar, free_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (a.T + WGH_ROWS - 1) / WGH_ROWS;
  const long long n_units = (long long)a.L * a.B * nkb;
  const long long per = (n_units + gridDim.x - 1) / gridDim.x;
  const long long u0 = (long long)blockIdx.x * per;
  long long u1 = u0 + per;
  if (u1 > n_units) u1 = n_units;
  if (u0 >= u1) return;

  for (int i = tid; i < WGH_ROWS * 64; i += blockDim.x) {
    const int rr = i / 64, cc = i % 64;
    *reinterpret_cast<__half*>(Ones + swzh(rr, cc)) = __float2half_rn(cc == 0 ? 1.f : 0.f);
  }
  for (int i = tid; i < STG * (int)(WGH_BLK / 16); i += blockDim.x) {      // dx' blocks: defined (zero) before the first load
    const int s = i / (WGH_BLK / 16), o = i % (WGH_BLK / 16);
    *reinterpret_cast<uint4*>(smem + s * WGH_STAGE + 4 * WGH_BLK + o * 16) = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < STG; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    mbar_init(&free_bar, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;      // columns 0-63 D1, 64-191 D2
  const int per_lb = nkb;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t i = 0;
      for (long long u = u0; u < u1; ++u, ++i) {
        const int lb = (int)(u / per_lb), kb = (int)(u - (long long)lb * per_lb);
        const int l = lb / a.B, b = lb - l * a.B;
        const bool hd = (l < a.L - 1) || a.last_dense;
        const int s = i % STG;
        const uint32_t ph = (i / STG) & 1;
        const int t = kb * WGH_ROWS;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], (hd ? 5 : 4) * WGH_BLK);
        unsigned char* sa = smem + s * WGH_STAGE;
        tma_load_3d(sa, &mapX, &full_bar[s], 0, t, lb);                          // x[t] of layer l (hi | lo)
        tma_load_3d(sa + WGH_BLK, &mapX, &full_bar[s], 0, t - a.dil[l], lb);     // x[t-d]  (zeros for t < d)
        tma_load_3d(sa + 2 * WGH_BLK, &mapZ, &full_bar[s], l * C, t, b);         // z_l | z_l+1 (columns past the end: zeros)
        tma_load_3d(sa + 3 * WGH_BLK, &mapP, &full_bar[s], 0, t, lb);            // df | dg
        if (hd) tma_load_3d(sa + 4 * WGH_BLK, &mapDn, &full_bar[s], 0, t, lb + a.B);   // dx' = dx of layer l+1 (hi | lo)
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID1 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t ID2 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t i = 0, seg = 0;
      int cur_lb = -1;
      for (long long u = u0; u < u1; ++u, ++i) {
        const int lb = (int)(u / per_lb);
        const bool first = (lb != cur_lb);
        if (first) {
          if (cur_lb >= 0) {
            mma_commit(&done_bar);                       // segment finished: hand the accumulators to the epilogue ...
            mbar_wait(&free_bar, seg & 1);               // ... and wait until they have been read
            tc_fence_after();
            ++seg;
          }
          cur_lb = lb;
        }
        const int s = i % STG;
        const uint32_t ph = (i / STG) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * WGH_STAGE);
        const uint64_t dA1 = mn16_desc(sa, WGH_BLK);                                          // [x | x[t-d]]
        const uint64_t dA2 = mn16_desc(sa + 2 * WGH_BLK, smem_u32(Ones) - (sa + 2 * WGH_BLK));   // [z | ones]
        const uint64_t dB = mn16_desc(sa + 3 * WGH_BLK, WGH_BLK);                             // [dpre | dx']
#pragma unroll
        for (int k = 0; k < WGH_ROWS / 16; ++k) {      // +2048 B per K = 16 time steps
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          mma_f16_ss(tmem, dA1 + 128 * k, dB + 128 * k, ID1, acc);
          mma_f16_ss(tmem + 64, dA2 + 128 * k, dB + 128 * k, ID2, acc);
        }
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
      const bool hd = (l < a.L - 1) || a.last_dense;
      long long ue = (long long)(lb + 1) * per_lb;      // end of this (layer, batch element) inside the range
      if (ue > u1) ue = u1;
      mbar_wait(&done_bar, seg & 1);
      tc_fence_after();
      const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
      {   // D1: rows = [x hi | x lo | x[t-d] hi | x[t-d] lo] channels, columns = df | dg
        const int tap = row < 64 ? 1 : 0, rch = row & 31;
        float* gf = a.gwf + (size_t)l * 2 * C * C + (size_t)(tap * C + rch) * C;
        float* gg = a.gwg + (size_t)l * 2 * C * C + (size_t)(tap * C + rch) * C;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_base + c0, v);   // warp-collective: every lane issues it
          float* dst = c0 == 0 ? gf : gg;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4h(dst + j, __uint_as_float(v[j]) * a.scale, __uint_as_float(v[j + 1]) * a.scale,
                        __uint_as_float(v[j + 2]) * a.scale, __uint_as_float(v[j + 3]) * a.scale);
        }
      }
      {   // D2: rows 0-31 z_l, row 64 ones; columns 0-63 dpre, 64-95 dx' hi, 96-127 dx' lo
        uint32_t v[32], w[32];
        if (quad == 2) {      // (warp-uniform) the ones row lives in this quadrant: bias / conditioning sums
          tmem_ld32(lane_base + 64, v);
          tmem_ld32(lane_base + 96, w);
          if (row == 64) {
            float* dst = a.gprebias + ((size_t)l * a.B + b) * 64;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              red_add_v4h(dst + j, __uint_as_float(v[j]) * a.scale, __uint_as_float(v[j + 1]) * a.scale,
                          __uint_as_float(v[j + 2]) * a.scale, __uint_as_float(v[j + 3]) * a.scale);
              red_add_v4h(dst + 32 + j, __uint_as_float(w[j]) * a.scale, __uint_as_float(w[j + 1]) * a.scale,
                          __uint_as_float(w[j + 2]) * a.scale, __uint_as_float(w[j + 3]) * a.scale);
            }
          }
        }
        if ((quad == 0 || quad == 2) && hd) {
          tmem_ld32(lane_base + 128, v);
          tmem_ld32(lane_base + 160, w);
          float* dst = nullptr;
          if (quad == 0) dst = a.gdense + (size_t)l * C * C + (size_t)row * C;      // dense[c][r], c = row
          else if (row == 64 && a.gdense_bias) dst = a.gdense_bias + (size_t)l * C;
          if (dst) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4h(dst + j, (__uint_as_float(v[j]) + __uint_as_float(w[j])) * a.scale,
                          (__uint_as_float(v[j + 1]) + __uint_as_float(w[j + 1])) * a.scale,
                          (__uint_as_float(v[j + 2]) + __uint_as_float(w[j + 2])) * a.scale,
                          (__uint_as_float(v[j + 3]) + __uint_as_float(w[j + 3])) * a.scale);
          }
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
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// xs / dxs / p16 as in block_bwd_chain; zcat16: [B][T][ldz] fp16 (ldz = L * 32 in the training step)
int block_wgrad_h_all(const void* xs, const void* dxs, const void* p16, const void* zcat16, int ldz, float scale, float* gwf,
                      float* gwg, float* gdense, float* gprebias, float* gdense_bias, const int* dilations, int L, int B,
                      int T, cudaStream_t st, int last_dense) {
  if (L < 1 || L > WN_MAX_LAYERS) return -1;
  CUtensorMap mX, mZ, mP, mDn;
  int rc = make_map_h64(&mX, xs, (int64_t)L * B, T, 64, 64, WGH_ROWS);
  if (rc) return rc;
  rc = make_map_h64(&mZ, zcat16, B, T, ldz, ldz, WGH_ROWS);
  if (rc) return rc;
  rc = make_map_h64(&mP, p16, (int64_t)L * B, T, 64, 64, WGH_ROWS);
  if (rc) return rc;
  rc = make_map_h64(&mDn, dxs, (int64_t)(L + (last_dense ? 1 : 0)) * B, T, 64, 64, WGH_ROWS);
  if (rc) return rc;
  WgHArgs a;
  a.gwf = gwf; a.gwg = gwg; a.gdense = gdense; a.gprebias = gprebias; a.gdense_bias = gdense_bias;
  a.L = L; a.B = B; a.T = T; a.last_dense = last_dense ? 1 : 0; a.scale = scale;
  for (int l = 0; l < WN_MAX_LAYERS; ++l) a.dil[l] = l < L ? dilations[l] : 0;
  const size_t smem = 1024 + WGH_STAGES * WGH_STAGE + WGH_BLK;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(block_wgrad_h_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const long long n_units = (long long)L * B * ((T + WGH_ROWS - 1) / WGH_ROWS);
  int grid = sm_count();
  if (grid > n_units) grid = (int)n_units;
  block_wgrad_h_all_kernel<<<grid, 192, smem, st>>>(mX, mZ, mP, mDn, a);
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_WGRAD);
  return 0;
}

int block_bwd_h_set_trap_info(unsigned int* p) { return umma::set_trap_info_tu(p); }

}  // namespace wn
