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
// Work items (PRE and DX of a tile are separate items, see the chain kernel below) are claimed from one global counter.
//   warps 0-7  epilogue (thread = one time step x 16 channels)
//   warp  8    loader: claims work, polls flags, TMA loads, weight images
//   warp  9    publisher: TMA stores of dpre and dx, then (stores complete) the tile's flags
//   warp 10    MMA issuer: tcgen05 MMAs of an item as soon as its operand tiles have landed
// (tcgen05 finding kept for the record, probed with the first, fused version of this kernel: a K-major 128B-swizzled
// operand may start at ANY row of a tile -- descriptor start address = tile + row * 128, matrix base offset field 0 --
// because the swizzle is a function of the absolute shared-memory address bits, not of the row index relative to the
// start address; with the base offset set to row & 7 the rows come out permuted.)
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
// per-layer weight image: W0h [64 n][32 k] | W1h (fp16, 64-byte rows, 64B swizzle: the recompute of the pre-activations is a
// single fp16 product, like the TF32 recompute of the first-generation kernels) | WdT [32 c][hi r | lo r] | Bcur [32 r][df c | dg c] | Bpast
constexpr uint32_t IMG_B = 5 * 4096;
constexpr uint32_t IMG_PRE_BYTES = 3 * 4096, IMG_DX_OFF = 3 * 4096, IMG_DX_BYTES = 2 * 4096;
constexpr uint32_t XH_TILE = TM * 64;         // [128 rows][32 halfs]
constexpr int BC_THREADS = 352;      // 8 epilogue warps + loader + publisher + MMA issuer

// byte offset of fp16 element (row, k) in a [rows][32 fp16] 64B-swizzled K-major tile
__device__ __forceinline__ uint32_t swzh64(int row, int k) {
  return (uint32_t)(row * 64 + ((((k >> 3) ^ ((row >> 1) & 3))) << 4) + ((k & 7) << 1));
}
// K-major operand, 64B swizzle: rows are 64 bytes, 8-row groups 512 bytes apart (SBO), layout type 4
__device__ __forceinline__ uint64_t kmajor_desc64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// the hi halves of split rows [B][T][hi 32 | lo 32]: box = [TM rows][32 halfs], 64B swizzle
static int make_map_xhi(CUtensorMap* m, const __half* ptr, int64_t B, int64_t T) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[3] = {32, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstr[2] = {128, (cuuint64_t)T * 128};
  cuuint32_t box[3] = {32, (cuuint32_t)TM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}
// two floats -> packed fp16 pair, saturating at +-65504 (one F2FP.SATFINITE instruction); lo = the lower half
__device__ __forceinline__ uint32_t pack_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// tanh(f) and sigmoid(g) for the backward recompute.  WN_BWD_TANH_APPROX: two MUFU.TANH (relative error 2^-11, the
// precision of the fp16 operands of the recompute itself) instead of two ex2 + one rcp + the range clamps.
#ifndef WN_BWD_TANH_APPROX
#define WN_BWD_TANH_APPROX 1
#endif
__device__ __forceinline__ void gate_parts(float f, float g, float& tf, float& sg) {
#if WN_BWD_TANH_APPROX
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(tf) : "f"(f));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * g));
  sg = fmaf(0.5f, t, 0.5f);
#else
  gated_parts_fast(f, g, tf, sg);
#endif
}
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

// ---- backward weight images: per layer  W0h | W1h | WdT | Bcur | Bpast  (fp16, swizzled K-major rows) ----
__global__ void block_bwd_h_images_kernel(unsigned char* __restrict__ img, const float* __restrict__ filter,
                                          const float* __restrict__ gate, const float* __restrict__ dense) {
  const int l = blockIdx.x;
  unsigned char* base = img + (size_t)l * IMG_B;
  const float* wf = filter + (size_t)l * 2 * C * C;
  const float* wg = gate + (size_t)l * 2 * C * C;
  const float* wd = dense + (size_t)l * C * C;
  for (int i = threadIdx.x; i < 2 * 64 * 32; i += blockDim.x) {  // W[tap]h: n = [filter | gate] output channel (fastest: coalesced), k = input channel
    const int tap = i / (64 * 32), k = (i / 64) % 32, n = i % 64;
    const float w = n < C ? wf[(tap * C + k) * C + n] : wg[(tap * C + k) * C + (n - C)];
    *reinterpret_cast<__half*>(base + tap * 4096 + swzh64(n, k)) = __float2half_rn(w);
  }
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {      // WdT: n = dilation channel c, k = residual channel r
    const int n = i / 32, k = i % 32;
    __half h, lo;
    split_h(wd[n * C + k], h, lo);
    *reinterpret_cast<__half*>(base + 8192 + swzh(n, k)) = h;
    *reinterpret_cast<__half*>(base + 8192 + swzh(n, 32 + k)) = lo;
  }
  for (int i = threadIdx.x; i < 2 * 32 * 64; i += blockDim.x) {  // B[tap]: n = residual channel r, k = [df c | dg c]
    const int which = i / (32 * 64), n = (i / 64) % 32, k = i % 64;
    const int tap = which == 0 ? 1 : 0;      // Bcur multiplies dpre[t] (tap 1 = current sample), Bpast dpre[t+d] (tap 0)
    const float w = k < 32 ? wf[(tap * C + n) * C + k] : wg[(tap * C + n) * C + (k - 32)];
    *reinterpret_cast<__half*>(base + IMG_DX_OFF + which * 4096 + swzh(n, k)) = __float2half_rn(w);
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
// Work items, claimed from one global counter, phase q = 0 .. L (n_tiles = B * ceil(T/128) items each):
//   q = 0        PRE(L-1, tile)                       pre-activation gradient of the last layer (it has no output gradient)
//   0 < q < L    DX(l, tile) + PRE(l-1, tile), l = L-q   input gradient of layer l, which IS the output gradient of layer
//                                                     l-1 at the same time steps: it stays in shared memory as the A
//                                                     operand of dx'.Wd^T and as the residual term of the next phase
//   q = L        DX(0, tile)
//   PRE(l): recompute pre-activations, dz = dz_skip + dx'.Wd^T, dpre -> P16[l]
//   DX(l) : dx = dx' + dpre[t].W1^T + dpre[t+d].W0^T -> DXS[l]          (dpre of the tile and of the tiles d steps later)
// An item only depends on items of the previous phase (deadlock-free for any number of resident CTAs), i.e. on items
// >= n_tiles - 5 back in the order.  History of this structure (tools/timeline_bwd.py):
//   * PRE and DX of the SAME layer fused per tile: a tile needs the dpre rows of the tile claimed right before it, the
//     two run in lockstep on different CTAs and every tile stalled ~5,000 cycles on that flag;
//   * PRE and DX as separate items (100 phases): no lockstep, but twice the dependent hops -- with 782 tiles per phase
//     and 296 CTAs a phase lasts about as long as one item's latency (claim -> loads -> MMA -> epilogue -> store -> flag),
//     so half of the items still found their flags unset;
//   * this form: 51 phases, one L2 round trip per layer instead of two, dx never re-loaded by its own consumer.
// Per CTA the items are pipelined: TMEM accumulators double buffered by item parity, a loader warp that prefetches the
// next item's tiles as soon as the MMAs have read the current ones, a separate MMA-issuing warp, a publisher warp.
enum { BC_W8 = 8, BC_W9 = 9, BC_W10 = 10 };

struct BwdChainArgs {
  const unsigned char* img_f;      // [L] forward weight images (IMG_H bytes each; W0cat | W1cat are used)
  const unsigned char* img_b;      // [L] backward weight images (IMG_B bytes each)
  const float* prebias;            // [L][B][64]
  unsigned int* flags;             // [L][n_tiles] dpre published | [L][n_tiles] dx published | work counter; zeroed before the launch
  int L, B, T, n_tt;
  int last_dense;                  // the last layer has an output gradient too (stand-alone wn_block_bwd)
  int late;                        // (experiment) reload Xh / Ob / Dz only after the PRE epilogue of the previous item
  float cs;                        // the skip-path gradient is multiplied by cs on its way into the chain's scaled domain
  long long* timeline;             // debug: cycles CTA 0's warps spent in each kind of wait
  int dil[WN_MAX_LAYERS];
};
static long long* g_timeline_b = nullptr;
void set_bwd_h_timeline(long long* p) { g_timeline_b = p; }

// Bounded waits with role-specific bounds (loader / MMA issuer < epilogue < publisher): when the protocol hangs, the
// trap info names what the LOADER was waiting for (tag = source line), which is what decides everything else.
__device__ __forceinline__ void mbar_wait_n(uint64_t* b, uint32_t parity, uint32_t log2n, uint32_t tag) {
  for (uint32_t i = 0; i < (1u << log2n); ++i) {
    uint32_t ok;
#ifdef WN_BWD_WAIT_HINT
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity), "r"((uint32_t)WN_BWD_WAIT_HINT) : "memory");
#else
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
#endif
    if (ok) return;
  }
  wait_timed_out(smem_u32(b), parity, tag, gridDim.y);
}
__device__ __forceinline__ void spin_until(volatile int* cnt, int need, uint32_t tag) {
  uint32_t spin = 0;
#pragma unroll 1
  while (*cnt < need) {
    if (++spin > (1u << 22)) wait_timed_out(0x5F4Eu, (unsigned int)need, tag, (unsigned int)*cnt);
  }
}
// single-lane roles (loader, MMA issuer, publisher): poll without suspending first -- their wake-up latency is on the
// critical path of every item, and one polling lane costs next to nothing (WN_BWD_SPIN polls before the suspending waits)
#ifndef WN_BWD_SPIN
#define WN_BWD_SPIN 0
#endif
__device__ __forceinline__ void mbar_spin_n(uint64_t* b, uint32_t parity, uint32_t log2n, uint32_t tag) {
#pragma unroll 1
  for (int i = 0; i < WN_BWD_SPIN; ++i)
    if (mbar_test(b, parity)) return;
  mbar_wait_n(b, parity, log2n, tag);
}
#define IWAIT(bar, par) mbar_spin_n(bar, par, 20, __LINE__)
#define EWAIT(bar, par) mbar_wait_n(bar, par, 22, __LINE__)
#define PWAIT(bar, par) mbar_spin_n(bar, par, 24, __LINE__)

__global__ void __launch_bounds__(BC_THREADS, 2)
block_bwd_chain_kernel(const __grid_constant__ CUtensorMap mapXH, const __grid_constant__ CUtensorMap mapDX,
                       const __grid_constant__ CUtensorMap mapDz, const __grid_constant__ CUtensorMap mapP,
                       const __grid_constant__ BwdChainArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* X0 = smem;                 // dpre[t]      } operands of the DX products
  unsigned char* X1 = smem + TILE;          // dpre[t+d]    }
  unsigned char* Ob = smem + 2 * TILE;      // split rows dx' in -> dx out (in place): TMA store source AND A operand of dx.Wd^T
  unsigned char* Pb = smem + 3 * TILE;      // dpre out [128][df 32 | dg 32]
  unsigned char* Xh0 = smem + 4 * TILE;     // hi halves of x[t]   [128][32], 64B swizzle   } operands of the pre-activation recompute
  unsigned char* Xh1 = Xh0 + XH_TILE;       // hi halves of x[t-d]                          }
  unsigned char* Dz = Xh1 + XH_TILE;        // skip-path gradient [128][32 halfs], 64B swizzle
  unsigned char* W0 = Dz + TM * 64;         // [64][32] past tap (filter | gate), 64B swizzle    } image of the PRE layer
  unsigned char* W1 = W0 + 4096;            // current tap                                       }
  unsigned char* WdT = W1 + 4096;           // [32 c][hi r | lo r]                               }
  unsigned char* Bc = WdT + 4096;           // [32 r][df c | dg c], tap 1                 } image of the DX layer
  unsigned char* Bp = Bc + 4096;            // tap 0                                      }
  // Every barrier completes exactly once per item (parts an item does not have arrive without data): phase parity =
  // item parity.  The accumulator / staging barriers exist once per item parity: an mbarrier whose phase is not looked
  // at before it completes twice more aliases, and the MMAs / epilogue of item i+1 may finish before a slow waiter has
  // looked at item i.
  __shared__ __align__(8) uint64_t bar_xa, bar_da, bar_xb, bar_m1[2], bar_mx[2], bar_m2[2], bar_zr, bar_o1[2], bar_o2[2], bar_sfree, bar_w;
  __shared__ uint32_t tmem_slot;
  __shared__ int item_s[6];                 // work item of tile i in item_s[i & 3] (-1: no more work), published through bar_xa / bar_da
  __shared__ __align__(16) float pb_s[64];
  // items whose dx store has read Ob (it may be re-loaded once the dx.Wd^T product has read it too): a monotonic counter,
  // because the loader does not need it every time (see above)
  __shared__ volatile int ob_cnt;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    ob_cnt = 0;
    mbar_init(&bar_xa, 1);
    mbar_init(&bar_da, 1);
    mbar_init(&bar_xb, 1);
    for (int k = 0; k < 2; ++k) {
      mbar_init(&bar_m1[k], 1);
      mbar_init(&bar_mx[k], 1);
      mbar_init(&bar_m2[k], 1);
      mbar_init(&bar_o1[k], 256);
      mbar_init(&bar_o2[k], 256);
    }
    mbar_init(&bar_zr, 256);
    mbar_init(&bar_sfree, 1);
    mbar_init(&bar_w, 1);
    mbar_fence_init();
  }
  if (warp == BC_W8) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // accumulators of item i at column 128 * (i & 1): 0-63 pre-activations (f | g), 64-95 dx.Wd^T, 96-127 the dx products
  const uint32_t tmem = tmem_slot;
  const int n_tiles = a.B * a.n_tt;
  const int n_items = (a.L + 1) * n_tiles;
  unsigned int* flagP = a.flags;
  unsigned int* flagX = a.flags + (size_t)a.L * n_tiles;
  // item -> (ld: layer of the DX part or -1, lp: layer of the PRE part or -1, batch element, tile)
  auto decode = [&](int item, int& ld, int& lp, int& b, int& tt) {
    const int q = item / n_tiles, j = item - q * n_tiles;
    ld = q >= 1 ? a.L - q : -1;
    lp = q <= a.L - 1 ? a.L - 1 - q : -1;
    b = j / a.n_tt;
    tt = j - b * a.n_tt;
  };
  // does the PRE part / DX part of layer l have an output gradient dx' (the last layer of the network has none)
  auto has_dn = [&](int l) { return l < a.L - 1 || a.last_dense; };

  if (warp == BC_W8) {
    // ------------------------------------------------------------------------------------------------ loader
    // The whole warp runs this loop (uniform control flow): lane 0 claims items and talks to the barriers, lanes 0-3 read
    // the (at most four) flags of an item and issue its four operand loads, lanes 0-1 the two epilogue-facing loads --
    // one thread issuing six TMA loads and polling four flags one after the other took ~2,000 cycles per item, on the
    // critical path between "item claimed" and "tiles landed".
    {
      const unsigned FULL = 0xffffffffu;
      unsigned int* counter = a.flags + 2 * (size_t)a.L * n_tiles;
      auto claim = [&]() -> int {
        unsigned int w = 0;
        if (lane == 0) w = atomicAdd(counter, 1u);
        w = __shfl_sync(FULL, w, 0);
        return w < (unsigned int)n_items ? (int)w : -1;
      };
      auto lwait = [&](uint64_t* bar, uint32_t parity, uint32_t tag) {      // lane 0 polls, the warp follows
        if (lane == 0) mbar_wait_n(bar, parity, 20, tag);
        __syncwarp();
      };
      // weights of an item: Bc | Bp of the DX layer, W0 | W1 | WdT of the PRE layer (the same for every item of a phase)
      auto load_weights = [&](int ld, int lp) {
        if (lane == 0) {
          mbar_expect_tx(&bar_w, (ld >= 0 ? IMG_DX_BYTES : 0u) + (lp >= 0 ? IMG_PRE_BYTES : 0u));
          if (lp >= 0) bulk_g2s(W0, a.img_b + (size_t)lp * IMG_B, IMG_PRE_BYTES, &bar_w);
          if (ld >= 0) bulk_g2s(Bc, a.img_b + (size_t)ld * IMG_B + IMG_DX_OFF, IMG_DX_BYTES, &bar_w);
        }
      };
      long long c_fl = 0, c_ob = 0, c_zr = 0, c_m1 = 0, c_m2 = 0, t_a, t_start = clock64();
      long long n_slow = 0;
      // flags of an item: dpre of its tile and of the tile(s) d steps later, dx of the layer above -- the previous phase.
      // Lane k holds flag k.  Relaxed loads: an acquire load holds back every later memory operation of the thread until
      // it has returned; the consumer of a set flag is a TMA load issued behind a branch on the value, and it reads the
      // L2, where the producer's stores had completed before it released the flag.
      const unsigned int* fmine = nullptr;
      unsigned int fval = 1u;
      auto flags_of = [&](int ld, int lp, int b, int tt) {
        fmine = nullptr;
        if (ld >= 0) {
          const unsigned int* f = flagP + (size_t)ld * n_tiles + (size_t)b * a.n_tt;
          const int t0 = tt * TM, d = a.dil[ld];
          const int ta = (t0 + d) / TM, tb = (t0 + d + TM - 1) / TM;
          if (lane == 0) fmine = f + tt;
          if (lane == 1 && ta < a.n_tt && ta != tt) fmine = f + ta;
          if (lane == 2 && tb < a.n_tt && tb != ta) fmine = f + tb;
          if (lane == 3 && ld < a.L - 1) fmine = flagX + (size_t)(ld + 1) * n_tiles + (size_t)b * a.n_tt + tt;
        }
        fval = fmine ? ld_relaxed(fmine) : 1u;
      };
      auto flags_wait = [&]() {
        if (__all_sync(FULL, fval != 0u)) return;
        t_a = clock64();
        fval = fmine ? ld_relaxed(fmine) : 1u;      // one fresh look before the acquire-polling slow path
        if (!__all_sync(FULL, fval != 0u)) {
          ++n_slow;
          if (fmine && !fval) wait_flag(fmine);
          __syncwarp();
        }
        c_fl += clock64() - t_a;
      };
      // operand tiles of an item: dpre[t], dpre[t+d] for the DX products; hi(x[t]), hi(x[t-d]) for the recompute.
      // (arrive / expect_tx = release: publishes item_s to the waiters of these phases)
      auto load_ops_a = [&](int ld, int lp, int b, int tt) {
        const int t0 = tt * TM;
        if (lane == 0) {
          if (ld >= 0) mbar_expect_tx(&bar_xa, 2 * TILE);
          else mbar_arrive(&bar_xa);
        }
        __syncwarp();
        if (ld >= 0 && lane < 2)      // (rows at or past the window end arrive as zeros)
          tma_load_3d(lane == 0 ? X0 : X1, &mapP, &bar_xa, 0, lane == 0 ? t0 : t0 + a.dil[ld], ld * a.B + b);
      };
      auto load_ops_b = [&](int ld, int lp, int b, int tt) {
        const int t0 = tt * TM;
        if (lane == 0) {
          if (lp >= 0) mbar_expect_tx(&bar_xb, 2 * XH_TILE);
          else mbar_arrive(&bar_xb);
        }
        __syncwarp();
        if (lp >= 0 && (lane == 2 || lane == 3))      // (rows before the window start arrive as zeros)
          tma_load_3d(lane == 2 ? Xh0 : Xh1, &mapXH, &bar_xb, 0, lane == 2 ? t0 : t0 - a.dil[lp], lp * a.B + b);
      };
      auto load_ops = [&](int ld, int lp, int b, int tt) {
        load_ops_a(ld, lp, b, tt);
        load_ops_b(ld, lp, b, tt);
      };
      // the tiles the epilogue threads read: dx' (the input gradient of the layer above) into Ob, the skip-path gradient into Dz
      auto load_dn = [&](int ld, int lp, int b, int tt) {
        const int lsrc = ld >= 0 ? ld : lp;
        const bool hd = has_dn(lsrc);
        const uint32_t bytes = (hd ? TILE : 0u) + (lp >= 0 ? (uint32_t)(TM * 64) : 0u);
        if (lane == 0) {
          if (bytes == 0) mbar_arrive(&bar_da);
          else mbar_expect_tx(&bar_da, bytes);
        }
        __syncwarp();
        if (hd && lane == 0) tma_load_3d(Ob, &mapDX, &bar_da, 0, tt * TM, (lsrc + 1) * a.B + b);
        if (lp >= 0 && lane == 1) tma_load_3d(Dz, &mapDz, &bar_da, lp * C, tt * TM, b);
      };
      int item = claim();
      if (lane == 0) item_s[0] = item;
      uint32_t i = 0;
      int ld = -1, lp = -1, b = 0, tt = 0, wq = -1;
      int nx = -1, nld = -1, nlp = -1, nb = 0, ntt = 0;
      if (item < 0) {
        if (lane == 0) {
          mbar_arrive(&bar_xa);
          mbar_arrive(&bar_da);
        }
      } else {
        decode(item, ld, lp, b, tt);
        load_weights(ld, lp);
        wq = item / n_tiles;
        flags_of(ld, lp, b, tt);
        flags_wait();
        load_ops(ld, lp, b, tt);
        load_dn(ld, lp, b, tt);
      }
      while (item >= 0) {
        const uint32_t par = i & 1, ph = (i >> 1) & 1;
        // every epilogue thread has read Dz of item i: it has left item i-1 (the accumulators of the other parity are free
        // for the MMAs of item i+1) and is past its wait on bar_da of item i
        t_a = clock64();
        lwait(&bar_zr, par, __LINE__);
        c_zr += clock64() - t_a;
        // Item i+1 is claimed as LATE as possible (the PRE epilogue of item i has started): with 782 tiles per phase and
        // 296 CTAs, every claimed-but-unstarted item shortens the distance (in time) to the items of the previous phase it
        // depends on.  Claimed here, their flags are practically always set; claimed one DX epilogue earlier (measured),
        // a quarter of the items found a flag unset and the stalls fed on each other: 11,300 instead of 9,600 cycles per item.
        nx = claim();
        if (nx >= 0) {
          decode(nx, nld, nlp, nb, ntt);
          flags_of(nld, nlp, nb, ntt);
        }
        t_a = clock64();
        lwait(&bar_mx[par], ph, __LINE__);      // the DX products and the recompute have read X0 / X1 / Xh0 / Xh1
        c_m1 += clock64() - t_a;
        if (lane == 0) item_s[(i + 1) & 3] = nx;
        __syncwarp();
        if (nx >= 0 && nx / n_tiles == wq) {      // (a new phase's weights: below, once every MMA of item i has completed)
          flags_wait();
          load_ops_a(nld, nlp, nb, ntt);
          if (!a.late) load_ops_b(nld, nlp, nb, ntt);
        }
        t_a = clock64();
        lwait(&bar_m2[par], ph, __LINE__);      // dx.Wd^T has read Ob (and every MMA of item i its weights)
        c_m2 += clock64() - t_a;
        if (nx < 0) {
          if (lane == 0) {
            mbar_arrive(&bar_xa);       // complete the next phases without data: the other warps see "no more work"
            mbar_arrive(&bar_da);
          }
          break;
        }
        if (nx / n_tiles != wq) {     // (the MMA warp waits for bar_w when it meets the first item of a phase)
          load_weights(nld, nlp);
          wq = nx / n_tiles;
          flags_wait();
          load_ops_a(nld, nlp, nb, ntt);
          if (!a.late) load_ops_b(nld, nlp, nb, ntt);
        }
        if (a.late) {
          lwait(&bar_o2[par], ph, __LINE__);
          load_ops_b(nld, nlp, nb, ntt);
        }
        if (ld >= 0) {                // the dx store of item i has left Ob
          t_a = clock64();
          if (lane == 0 && ob_cnt < (int)i + 1) spin_until(&ob_cnt, (int)i + 1, __LINE__);
          __syncwarp();
          c_ob += clock64() - t_a;
        }
        load_dn(nld, nlp, nb, ntt);
        item = nx; ld = nld; lp = nlp; b = nb; tt = ntt;
        ++i;
      }
      if (a.timeline && blockIdx.x == 0 && lane == 0) {
        a.timeline[0] = clock64() - t_start; a.timeline[3] = c_fl; a.timeline[4] = c_ob; a.timeline[5] = c_zr; a.timeline[6] = c_m1;
        a.timeline[8] = c_m2; a.timeline[9] = i; a.timeline[10] = gridDim.x; a.timeline[21] = n_slow;
      }
    }
  } else if (warp == BC_W10) {
    // ------------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t ID64 = idesc_f16(128, 64), ID32 = idesc_f16(128, 32);
      const uint64_t dX0 = kmajor_desc(smem_u32(X0)), dX1 = kmajor_desc(smem_u32(X1)), dOb = kmajor_desc(smem_u32(Ob));
      const uint64_t dXh0 = kmajor_desc64(smem_u32(Xh0)), dXh1 = kmajor_desc64(smem_u32(Xh1));
      const uint64_t dW0 = kmajor_desc64(smem_u32(W0)), dW1 = kmajor_desc64(smem_u32(W1)), dWdT = kmajor_desc(smem_u32(WdT));
      const uint64_t dBc = kmajor_desc(smem_u32(Bc)), dBp = kmajor_desc(smem_u32(Bp));
      long long c_xa = 0, c_xb = 0, c_o1 = 0, c_w = 0, b_mma = 0, t_a;
      int wq = -1;
      uint32_t wphase = 0;
      for (uint32_t i = 0;; ++i) {
        const uint32_t par = i & 1;
        const uint32_t acc = tmem + 128 * par;      // (free: the loader armed bar_xa only after every epilogue thread had left item i-2)
        t_a = clock64();
        IWAIT(&bar_xa, par);
        c_xa += clock64() - t_a;
        const int item = item_s[i & 3];
        if (item < 0) break;
        int ld, lp, b, tt;
        decode(item, ld, lp, b, tt);
        if (item / n_tiles != wq) {      // first item of a phase: its weight images
          t_a = clock64();
          IWAIT(&bar_w, wphase);
          c_w += clock64() - t_a;
          wphase ^= 1;
          wq = item / n_tiles;
        }
        tc_fence_after();
        t_a = clock64();
        if (ld >= 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16_ss(acc + 96, dX0 + 2 * k, dBc + 2 * k, ID32, k > 0);      // dpre[t]   . W[1]^T
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16_ss(acc + 96, dX1 + 2 * k, dBp + 2 * k, ID32, 1u);         // dpre[t+d] . W[0]^T
          mma_commit(&bar_m1[par]);
        } else {
          mbar_arrive(&bar_m1[par]);
        }
        b_mma += clock64() - t_a;
        if (lp >= 0) {
          t_a = clock64();
          IWAIT(&bar_xb, par);
          c_xb += clock64() - t_a;
          tc_fence_after();
          t_a = clock64();
          mma_f16_ss(acc, dXh1 + 0, dW0 + 0, ID64, 0u);       // hi(x[t-d]) . W[0]
          mma_f16_ss(acc, dXh1 + 2, dW0 + 2, ID64, 1u);
          mma_f16_ss(acc, dXh0 + 0, dW1 + 0, ID64, 1u);       // hi(x[t])   . W[1]
          mma_f16_ss(acc, dXh0 + 2, dW1 + 2, ID64, 1u);
          mma_commit(&bar_mx[par]);
          b_mma += clock64() - t_a;
          if (has_dn(lp)) {      // the output gradient of layer lp: this item's dx (staged by the epilogue threads), or the loaded dx'
            t_a = clock64();
            if (ld >= 0) IWAIT(&bar_o1[par], (i >> 1) & 1);
            else IWAIT(&bar_da, par);
            c_o1 += clock64() - t_a;
            tc_fence_after();
            t_a = clock64();
            mma_split(acc + 64, dOb, dWdT, ID32, true);      // dx' . Wd^T
            b_mma += clock64() - t_a;
          }
          mma_commit(&bar_m2[par]);
        } else {
          mma_commit(&bar_mx[par]);
          mma_commit(&bar_m2[par]);
        }
      }
      if (a.timeline && blockIdx.x == 0) {
        a.timeline[1] = c_xa; a.timeline[2] = c_xb; a.timeline[7] = c_w; a.timeline[20] = b_mma; a.timeline[22] = c_o1;
      }
    }
  } else if (warp == BC_W9) {
    // ------------------------------------------------------------------------------------------------ publisher
    if (lane == 0) {
      for (uint32_t i = 0;; ++i) {
        const uint32_t par = i & 1, ph = (i >> 1) & 1;
        PWAIT(&bar_o1[par], ph);
        const int item = item_s[i & 3];
        if (item < 0) break;
        int ld, lp, b, tt;
        decode(item, ld, lp, b, tt);
        const size_t tile = (size_t)b * a.n_tt + tt;
        // rows past the end of the window are clipped by the tensor maps
        if (ld >= 0) {
          tma_store_3d(&mapDX, Ob, 0, tt * TM, ld * a.B + b);
          bulk_commit();
          bulk_wait_read0();
          ob_cnt = (int)i + 1;        // Ob has been read (the loader may refill it once dx.Wd^T has read it too)
          bulk_wait0();               // the store has completed and is visible to this thread ...
          __threadfence();
          st_release(flagX + (size_t)ld * n_tiles + tile, 1u);      // ... publish the tile's dx
        }
        PWAIT(&bar_o2[par], ph);
        if (lp >= 0) {
          tma_store_3d(&mapP, Pb, 0, tt * TM, lp * a.B + b);
          bulk_commit();
          bulk_wait_read0();
        }
        mbar_arrive(&bar_sfree);      // Pb may be rewritten (and: this thread has looked at both staging barriers of item i)
        if (lp >= 0) {
          bulk_wait0();
          __threadfence();
          st_release(flagP + (size_t)lp * n_tiles + tile, 1u);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------------ epilogue
    const int r = tid & 127, half = tid >> 7;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16 * half;
    const uint32_t row_off = (uint32_t)r * 128;
    const uint32_t ch0 = (uint32_t)((2 * half) ^ (r & 7)) << 4, ch1 = (uint32_t)((2 * half + 1) ^ (r & 7)) << 4;
    const uint32_t cl0 = (uint32_t)((4 + 2 * half) ^ (r & 7)) << 4, cl1 = (uint32_t)((5 + 2 * half) ^ (r & 7)) << 4;
    int key = -1;
    const bool tl_on = a.timeline && blockIdx.x == 0 && tid == 0;
    long long e_da = 0, e_m1 = 0, e_xb = 0, e_m2 = 0, e_sf = 0, e_t, e_start = clock64();
    for (uint32_t i = 0;; ++i) {
      const uint32_t par = i & 1, ph = (i >> 1) & 1;
      const uint32_t lane_addr = lane_base + 128 * par;
      e_t = clock64();
      EWAIT(&bar_da, par);
      e_da += clock64() - e_t;
      const int item = item_s[i & 3];
      if (item < 0) {
        // hand the end marker on to the publisher (it has looked at every staging phase up to item i-1 once bar_sfree of
        // item i-1 has completed; the barrier of this parity was last used by item i-2)
        if (i > 0) EWAIT(&bar_sfree, par ^ 1u);
        mbar_arrive(&bar_o1[par]);
        if (tl_on) { a.timeline[11] = clock64() - e_start; a.timeline[12] = e_da; a.timeline[13] = e_m1; a.timeline[14] = e_xb; a.timeline[15] = e_m2; a.timeline[23] = e_sf; }
        break;
      }
      int ld, lp, b, tt;
      decode(item, ld, lp, b, tt);
      // ---- DX part: dx = dx' + the two products, in place in Ob ----
      if (ld >= 0) {
        float dx[16];
        if (has_dn(ld)) {
          const uint4 h0 = *reinterpret_cast<const uint4*>(Ob + row_off + ch0), h1 = *reinterpret_cast<const uint4*>(Ob + row_off + ch1);
          const uint4 l0 = *reinterpret_cast<const uint4*>(Ob + row_off + cl0), l1 = *reinterpret_cast<const uint4*>(Ob + row_off + cl1);
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
        e_t = clock64();
        EWAIT(&bar_m1[par], ph);
        e_m1 += clock64() - e_t;
        tc_fence_after();
        {
          uint32_t ov[16];
          tmem_ld16(lane_addr + 96, ov);
#pragma unroll
          for (int q = 0; q < 16; ++q) dx[q] += __uint_as_float(ov[q]);
        }
        uint32_t xh[8], xl[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {      // hi = fp16(dx) (saturating), lo = fp16(dx - hi)
          xh[q] = pack_sat(dx[2 * q], dx[2 * q + 1]);
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&xh[q]));
          xl[q] = pack_sat(dx[2 * q] - hf.x, dx[2 * q + 1] - hf.y);
        }
        *reinterpret_cast<uint4*>(Ob + row_off + ch0) = make_uint4(xh[0], xh[1], xh[2], xh[3]);
        *reinterpret_cast<uint4*>(Ob + row_off + ch1) = make_uint4(xh[4], xh[5], xh[6], xh[7]);
        *reinterpret_cast<uint4*>(Ob + row_off + cl0) = make_uint4(xl[0], xl[1], xl[2], xl[3]);
        *reinterpret_cast<uint4*>(Ob + row_off + cl1) = make_uint4(xl[4], xl[5], xl[6], xl[7]);
        fence_async_smem();
        tc_fence_before();
      }
      mbar_arrive(&bar_o1[par]);
      // ---- PRE part: dpre of the layer below (its output gradient is the dx just staged) ----
      if (lp >= 0) {
        if (lp * a.B + b != key) {      // uniform over the epilogue warps: bias / conditioning row of this (layer, batch element)
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (tid < 64) pb_s[tid] = a.prebias[((size_t)lp * a.B + b) * 64 + tid];
          key = lp * a.B + b;
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
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
        mbar_arrive(&bar_zr);
        e_t = clock64();
        EWAIT(&bar_m2[par], ph);
        e_m2 += clock64() - e_t;
        tc_fence_after();
        if (has_dn(lp)) {
          uint32_t av[16];
          tmem_ld16(lane_addr + 64, av);
#pragma unroll
          for (int q = 0; q < 16; ++q) dz[q] += __uint_as_float(av[q]);
        }
        if ((tt * TM + r) >= a.T) {      // rows past the window end are the (zero) future of the rows before them
#pragma unroll
          for (int q = 0; q < 16; ++q) dz[q] = 0.f;
        }
        uint32_t dfh[8], dgh[8];
        {
          uint32_t fv[16], gv[16];
          tmem_ld16(lane_addr + 0, fv);
          tmem_ld16(lane_addr + 32, gv);
          const float4* pbf = reinterpret_cast<const float4*>(pb_s + 16 * half);
          const float4* pbg = reinterpret_cast<const float4*>(pb_s + 32 + 16 * half);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 bf = pbf[q4], bg = pbg[q4];
            const float bfa[4] = {bf.x, bf.y, bf.z, bf.w}, bga[4] = {bg.x, bg.y, bg.z, bg.w};
            float df[4], dg[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int q = 4 * q4 + e;
              float tf, sg;
              gate_parts(__uint_as_float(fv[q]) + bfa[e], __uint_as_float(gv[q]) + bga[e], tf, sg);
              const float zs = dz[q] * sg;
              df[e] = zs * fmaf(-tf, tf, 1.f);
              dg[e] = zs * tf * (1.f - sg);
            }
            dfh[2 * q4] = pack_sat(df[0], df[1]); dfh[2 * q4 + 1] = pack_sat(df[2], df[3]);
            dgh[2 * q4] = pack_sat(dg[0], dg[1]); dgh[2 * q4 + 1] = pack_sat(dg[2], dg[3]);
          }
        }
        e_t = clock64();
        if (i > 0) EWAIT(&bar_sfree, par ^ 1u);      // the previous item's dpre store has left Pb
        e_sf += clock64() - e_t;
        *reinterpret_cast<uint4*>(Pb + row_off + ch0) = make_uint4(dfh[0], dfh[1], dfh[2], dfh[3]);
        *reinterpret_cast<uint4*>(Pb + row_off + ch1) = make_uint4(dfh[4], dfh[5], dfh[6], dfh[7]);
        *reinterpret_cast<uint4*>(Pb + row_off + cl0) = make_uint4(dgh[0], dgh[1], dgh[2], dgh[3]);
        *reinterpret_cast<uint4*>(Pb + row_off + cl1) = make_uint4(dgh[4], dgh[5], dgh[6], dgh[7]);
        fence_async_smem();
        tc_fence_before();
      } else {
        mbar_arrive(&bar_zr);
        if (i > 0) EWAIT(&bar_sfree, par ^ 1u);      // (every thread looks at every phase of this barrier)
      }
      mbar_arrive(&bar_o2[par]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BC_W8) tmem_dealloc(tmem, 256);
}

// xs: [L][B][T][hi 32 | lo 32] layer inputs (the forward chain's ring with L slots); dxs: [L (+1)][B][T][hi | lo] input
// gradients in the scaled domain (slot l+1 is dx' of layer l; slot L exists only with last_dense); p16: [L][B][T][64]
// dpre; dz16: [B][T][ldz] skip-path gradient; flags: 2 * L * B * ceil(T/128) + 1 words.
int block_bwd_chain(const void* xs, void* dxs, void* p16, const void* dz16, int ldz, float cs, const unsigned char* img_f,
                    const unsigned char* img_b, const float* prebias, const int* dilations, int L, int B, int T,
                    unsigned int* flags, cudaStream_t st, int last_dense) {
  if (L < 1 || L > WN_MAX_LAYERS) return -1;
  CUtensorMap mapXH, mapDX, mapDz, mapP;
  int rc = make_map_xhi(&mapXH, (const __half*)xs, (int64_t)L * B, T);
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
  a.last_dense = last_dense ? 1 : 0; a.cs = cs;
  { static const int late = [] { const char* e = getenv("WN_BWD_LATE"); return (e && e[0] == '1') ? 1 : 0; }(); a.late = late; }
  a.timeline = g_timeline_b;
  for (int l = 0; l < WN_MAX_LAYERS; ++l) a.dil[l] = l < L ? dilations[l] : 0;
  const size_t smem = 1024 + 4 * TILE + 2 * XH_TILE + TM * 64 + IMG_B;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(block_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -5;
    cudaFuncSetAttribute(block_bwd_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  const int64_t n_tiles = (int64_t)B * a.n_tt;
  if ((L + 1) * n_tiles >= (1ll << 30)) return -1;
  int grid = 2 * sm_count();      // two CTAs fit an SM (109 KB shared memory, 256 TMEM columns each)
  if (grid > n_tiles) grid = (int)n_tiles;
  cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)(2 * L * n_tiles + 1) * sizeof(unsigned int), st);
  if (e != cudaSuccess) return (int)e;
  block_bwd_chain_kernel<<<grid, BC_THREADS, smem, st>>>(mapXH, mapDX, mapDz, mapP, a);
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

static int block_wgrad_h32_all(const void* xs, const void* dxs, const void* p16, const void* zcat16, int ldz, float scale,
                               float* gwf, float* gwg, float* gdense, float* gprebias, float* gdense_bias,
                               const int* dilations, int L, int B, int T, cudaStream_t st, int last_dense);
// xs / dxs / p16 as in block_bwd_chain; zcat16: [B][T][ldz] fp16 (ldz = L * 32 in the training step)
int block_wgrad_h_all(const void* xs, const void* dxs, const void* p16, const void* zcat16, int ldz, float scale, float* gwf,
                      float* gwg, float* gdense, float* gprebias, float* gdense_bias, const int* dilations, int L, int B,
                      int T, cudaStream_t st, int last_dense) {
  if (L < 1 || L > WN_MAX_LAYERS) return -1;
  static const bool full = [] { const char* e = getenv("WN_WGRAD_FULL"); return e && e[0] == '1'; }();
  if (!full)
    return block_wgrad_h32_all(xs, dxs, p16, zcat16, ldz, scale, gwf, gwg, gdense, gprebias, gdense_bias, dilations, L, B, T,
                               st, last_dense);
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

// =====================================================================================================================
// The same weight gradients from HALF the bytes: only the hi halves of x and dx' are read (64 B of each 128 B row; dpre is
// fp16 anyway, so every product already carries an 11-bit factor) and z comes as a 32-column box instead of 64.  Per
// (layer, time step) the kernel reads 64 (x) + 64 (z) + 128 (dpre) + 64 (dx') = 320 B from HBM (x[t-d] hits L2) instead of
// 512.  All operands are MN-major blocks of [64 time steps][32 halfs] (64-byte rows, 64B swizzle, LBO = distance between
// blocks, SBO 512, 1024 B per K = 16 step) and ONE accumulator serves everything:
//   D[i][j] = sum_t [x | x[t-d] | z | 1,0..][t][i] * [df | dg | dx'][t][j]       (M = 128, N = 96)
//   lanes 0-31 x (tap 1), 32-63 x[t-d] (tap 0): columns 0-63 -> filter / gate gradients
//   lanes 64-95 z: columns 64-95 -> dense;   lane 96 (ones): columns 0-63 -> bias / conditioning sums, 64-95 -> dense bias
// =====================================================================================================================
constexpr int WG3_STAGES = 7;
constexpr uint32_t WG3_BLK = WGH_ROWS * 64;       // one [64 steps][32 halfs] block
constexpr uint32_t WG3_STAGE = 7 * WG3_BLK;       // x | x[t-d] | z | ones | df | dg | dx'

__device__ __forceinline__ uint64_t mn16_desc64(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// [rows][32 halfs] boxes (64B swizzle) out of rows of `ld` halfs; no L2 promotion: the other half of a 128 B line is not wanted
static int make_map_h32(CUtensorMap* m, const void* ptr, int64_t B, int64_t T, int64_t cols, int64_t ld, int rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}

__global__ void __launch_bounds__(192, 1)
block_wgrad_h32_all_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ,
                           const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapDn,
                           const __grid_constant__ WgHArgs a) {
  constexpr int STG = WG3_STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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

  // per stage: the constant block (column 0 = 1, i.e. the first half of every 64-byte row; the 64B swizzle moves 16-byte
  // chunk 0 of row r to chunk (r >> 1) & 3) and a defined (zero) dx' block
  for (int i = tid; i < STG * (int)(WG3_BLK / 16); i += blockDim.x) {
    const int s = i / (WG3_BLK / 16), o = i % (WG3_BLK / 16);
    const int rr = o >> 2, ch = o & 3;
    const bool one = ch == ((rr >> 1) & 3);
    *reinterpret_cast<uint4*>(smem + s * WG3_STAGE + 3 * WG3_BLK + o * 16) = make_uint4(one ? 0x3C00u : 0u, 0, 0, 0);
    *reinterpret_cast<uint4*>(smem + s * WG3_STAGE + 6 * WG3_BLK + o * 16) = make_uint4(0, 0, 0, 0);
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
        mbar_expect_tx(&full_bar[s], (hd ? 6 : 5) * WG3_BLK);
        unsigned char* sa = smem + s * WG3_STAGE;
        tma_load_3d(sa, &mapX, &full_bar[s], 0, t, lb);                          // x[t] of layer l (hi)
        tma_load_3d(sa + WG3_BLK, &mapX, &full_bar[s], 0, t - a.dil[l], lb);     // x[t-d]  (zeros for t < d)
        tma_load_3d(sa + 2 * WG3_BLK, &mapZ, &full_bar[s], l * C, t, b);         // z_l
        tma_load_3d(sa + 4 * WG3_BLK, &mapP, &full_bar[s], 0, t, lb);            // df
        tma_load_3d(sa + 5 * WG3_BLK, &mapP, &full_bar[s], 32, t, lb);           // dg
        if (hd) tma_load_3d(sa + 6 * WG3_BLK, &mapDn, &full_bar[s], 0, t, lb + a.B);   // dx' = dx of layer l+1 (hi)
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t ID = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t i = 0, seg = 0;
      int cur_lb = -1;
      for (long long u = u0; u < u1; ++u, ++i) {
        const int lb = (int)(u / per_lb);
        const bool first = (lb != cur_lb);
        if (first) {
          if (cur_lb >= 0) {
            mma_commit(&done_bar);
            mbar_wait(&free_bar, seg & 1);
            tc_fence_after();
            ++seg;
          }
          cur_lb = lb;
        }
        const int s = i % STG;
        const uint32_t ph = (i / STG) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * WG3_STAGE);
        const uint64_t dA = mn16_desc64(sa, WG3_BLK);
        const uint64_t dB = mn16_desc64(sa + 4 * WG3_BLK, WG3_BLK);
#pragma unroll
        for (int k = 0; k < WGH_ROWS / 16; ++k)      // +1024 B per K = 16 time steps
          mma_f16_ss(tmem, dA + 64 * k, dB + 64 * k, ID, (first && k == 0) ? 0u : 1u);
        mma_commit(&empty_bar[s]);
      }
      mma_commit(&done_bar);
    }
  } else {
    const int quad = warp & 3;
    uint32_t seg = 0;
    long long u = u0;
    while (u < u1) {
      const int lb = (int)(u / per_lb);
      const int l = lb / a.B, b = lb - l * a.B;
      const bool hd = (l < a.L - 1) || a.last_dense;
      long long ue = (long long)(lb + 1) * per_lb;
      if (ue > u1) ue = u1;
      mbar_wait(&done_bar, seg & 1);
      tc_fence_after();
      const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
      uint32_t v[32];
      if (quad < 2) {      // x (tap 1) / x[t-d] (tap 0): columns = df | dg
        const int tap = quad == 0 ? 1 : 0;
        float* gf = a.gwf + (size_t)l * 2 * C * C + (size_t)(tap * C + lane) * C;
        float* gg = a.gwg + (size_t)l * 2 * C * C + (size_t)(tap * C + lane) * C;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          tmem_ld32(lane_base + c0, v);
          float* dst = c0 == 0 ? gf : gg;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4h(dst + j, __uint_as_float(v[j]) * a.scale, __uint_as_float(v[j + 1]) * a.scale,
                        __uint_as_float(v[j + 2]) * a.scale, __uint_as_float(v[j + 3]) * a.scale);
        }
      } else if (quad == 2) {      // z: dense[c][r], c = lane
        if (hd) {
          tmem_ld32(lane_base + 64, v);
          float* dst = a.gdense + (size_t)l * C * C + (size_t)lane * C;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4h(dst + j, __uint_as_float(v[j]) * a.scale, __uint_as_float(v[j + 1]) * a.scale,
                        __uint_as_float(v[j + 2]) * a.scale, __uint_as_float(v[j + 3]) * a.scale);
        }
      } else {                     // the ones row (lane 0 of this quadrant): column sums
#pragma unroll 1
        for (int c0 = 0; c0 < 96; c0 += 32) {
          if (c0 == 64 && !(hd && a.gdense_bias)) break;      // (warp-uniform)
          tmem_ld32(lane_base + c0, v);
          if (lane == 0) {
            float* dst = c0 < 64 ? a.gprebias + ((size_t)l * a.B + b) * 64 + c0 : a.gdense_bias + (size_t)l * C;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4h(dst + j, __uint_as_float(v[j]) * a.scale, __uint_as_float(v[j + 1]) * a.scale,
                          __uint_as_float(v[j + 2]) * a.scale, __uint_as_float(v[j + 3]) * a.scale);
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
  if (warp == 1) tmem_dealloc(tmem, 128);
}

static int block_wgrad_h32_all(const void* xs, const void* dxs, const void* p16, const void* zcat16, int ldz, float scale,
                               float* gwf, float* gwg, float* gdense, float* gprebias, float* gdense_bias,
                               const int* dilations, int L, int B, int T, cudaStream_t st, int last_dense) {
  CUtensorMap mX, mZ, mP, mDn;
  int rc = make_map_h32(&mX, xs, (int64_t)L * B, T, 32, 64, WGH_ROWS);      // the hi halves only
  if (rc) return rc;
  rc = make_map_h32(&mZ, zcat16, B, T, ldz, ldz, WGH_ROWS);
  if (rc) return rc;
  rc = make_map_h32(&mP, p16, (int64_t)L * B, T, 64, 64, WGH_ROWS);
  if (rc) return rc;
  rc = make_map_h32(&mDn, dxs, (int64_t)(L + (last_dense ? 1 : 0)) * B, T, 32, 64, WGH_ROWS);
  if (rc) return rc;
  WgHArgs a;
  a.gwf = gwf; a.gwg = gwg; a.gdense = gdense; a.gprebias = gprebias; a.gdense_bias = gdense_bias;
  a.L = L; a.B = B; a.T = T; a.last_dense = last_dense ? 1 : 0; a.scale = scale;
  for (int l = 0; l < WN_MAX_LAYERS; ++l) a.dil[l] = l < L ? dilations[l] : 0;
  const size_t smem = 1024 + WG3_STAGES * WG3_STAGE;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(block_wgrad_h32_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const long long n_units = (long long)L * B * ((T + WGH_ROWS - 1) / WGH_ROWS);
  int grid = sm_count();
  if (grid > n_units) grid = (int)n_units;
  block_wgrad_h32_all_kernel<<<grid, 192, smem, st>>>(mX, mZ, mP, mDn, a);
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_WGRAD);
  return 0;
}

#include "block_bwd_fused.inc"

int block_bwd_h_set_trap_info(unsigned int* p) { return umma::set_trap_info_tu(p); }

}  // namespace wn
