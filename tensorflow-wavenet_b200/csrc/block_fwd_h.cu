// Forward gated residual block, second generation: fp16 split-precision operands on tcgen05 kind::f16.
// Reference: wavenet/model.py:236-330 (_create_dilation_layer), wavenet/ops.py:46-62 (causal_conv).
//
// The residual stream travels between layers as "split rows": every 32-channel fp32 row is stored as
// [hi(32 x fp16) | lo(32 x fp16)] = 128 bytes, hi = fp16(x), lo = fp16(x - hi)  (22 significant bits, the same
// as the hi/lo TF32 split of block_umma.cu, and the same bytes as the fp32 row).  A split row IS a K-major
// operand row of 64 fp16: K steps 0,1 are hi, K steps 2,3 are lo.  With weights stored the same way
// ([hi | lo] per output channel) a full-precision product is six K=16 MMAs per tap
//        hi.Whi (2)  +  lo.Whi (2)  +  hi.Wlo (2)
// selected purely by descriptor start addresses -- versus twelve K=8 TF32 MMAs plus a thread pass that splits
// the fp32 tile in block_umma.cu.  Per 128-step tile: no split pass, half the MMAs, half the shared-memory
// operand traffic.
// The kernel also writes the fp32 x' (kept for the backward pass) and z (tf32-rounded, into Zcat).
// Range: fp16 tops out at 65504; a residual stream that large has diverged anyway (documented in DESIGN.md).
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>

#include "../../include/wavenet_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "umma_common.cuh"
#include "split_common.cuh"

namespace wn {
using namespace umma;

// (tile constants, fp16 descriptors, split-row tensor maps, TMA stores, tile flags: split_common.cuh)
// ---- weight images: per layer  W0cat [64 n][hi 32 | lo 32] | W1cat | Wdcat [32 n][hi | lo]  (swizzled rows) ----
__global__ void block_h_images_kernel(unsigned char* __restrict__ img, const float* __restrict__ filter,
                                      const float* __restrict__ gate, const float* __restrict__ dense) {
  const int l = blockIdx.x;
  unsigned char* base = img + (size_t)l * IMG_H;
  const float* wf = filter + (size_t)l * 2 * C * C;
  const float* wg = gate + (size_t)l * 2 * C * C;
  const float* wd = dense + (size_t)l * C * C;
  // (one CTA per layer took 16 us on the step's critical path: 20 dependent iterations per thread; blockIdx.y splits them)
  const int tstride = blockDim.x * gridDim.y, tfirst = blockIdx.y * blockDim.x + threadIdx.x;
  for (int i = tfirst; i < 2 * 64 * 32; i += tstride) {      // tap, n (fastest: coalesced), k
    const int tap = i / (64 * 32), k = (i / 64) % 32, n = i % 64;
    const float w = n < C ? wf[(tap * C + k) * C + n] : wg[(tap * C + k) * C + (n - C)];
    __half h, lo;
    split_h(w, h, lo);
    unsigned char* t = base + tap * 8192;
    *reinterpret_cast<__half*>(t + swzh(n, k)) = h;
    *reinterpret_cast<__half*>(t + swzh(n, 32 + k)) = lo;
  }
  for (int i = tfirst; i < 32 * 32; i += tstride) {          // n = residual channel r, k = dilation channel d
    const int k = i / 32, n = i % 32;
    __half h, lo;
    split_h(wd[k * C + n], h, lo);
    unsigned char* t = base + 16384;
    *reinterpret_cast<__half*>(t + swzh(n, k)) = h;
    *reinterpret_cast<__half*>(t + swzh(n, 32 + k)) = lo;
  }
}

int64_t block_h_images_bytes(int L) { return (int64_t)L * IMG_H; }
uint32_t block_h_img_stride() { return IMG_H; }
int block_h_images(unsigned char* img, const float* filter, const float* gate, const float* dense, int L, cudaStream_t st) {
  block_h_images_kernel<<<dim3(L, 4), 256, 0, st>>>(img, filter, gate, dense);
  WN_CHECK_LAUNCH();
  return 0;
}

// fp32 rows [M][32] -> split rows [M][hi 32 | lo 32]   (input of the first layer)
__global__ void split_rows_kernel(const float* __restrict__ x, __half* __restrict__ xs, int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int64_t m = i >> 3;
    const int c4 = (int)(i & 7) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    __half h[4], l[4];
    split_h(v.x, h[0], l[0]); split_h(v.y, h[1], l[1]); split_h(v.z, h[2], l[2]); split_h(v.w, h[3], l[3]);
    __half* row = xs + m * 64;
    *reinterpret_cast<uint2*>(row + c4) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(row + 32 + c4) = *reinterpret_cast<uint2*>(l);
  }
}
int split_rows(const float* x, void* xs, int64_t M, cudaStream_t st) {
  const int64_t n4 = M * 8;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  split_rows_kernel<<<(int)blocks, 256, 0, st>>>(x, (__half*)xs, n4);
  WN_CHECK_LAUNCH();
  return 0;
}

struct FwdHArgs {
  const unsigned char* img;        // IMG_H bytes
  const float *prebias, *dense_bias;
  int B, T, d, is_last, zcol;      // zcol: first column of this layer inside Zcat
  int pdl_next;                    // the next kernel in the stream is launched programmatically and waits
  int z16;                         // also write z as fp16 (A operand of the fp16 skip GEMM)
  long long* timeline;             // debug (wn_debug_timeline): clock64 stamps of CTA 0, layers with d == 32 only
};
static long long* g_timeline_h = nullptr;
void set_fwd_h_timeline(long long* p) { g_timeline_h = p; }
#define TLH(i) do { if (a.timeline && blockIdx.x == 0 && tid == 0 && it < 4) a.timeline[it * 8 + (i)] = clock64(); } while (0)


// All three outputs of a tile (z -> Zcat column block, fp32 x', split x') leave through TMA stores from swizzled
// staging tiles: measured, per-thread 16-byte global stores (32 lines per warp instruction) cost 12 of the 22 us.
__global__ void __launch_bounds__(256, 2)
block_fwd_h_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ,
                   const __grid_constant__ CUtensorMap mapXo, const __grid_constant__ CUtensorMap mapXs,
                   const __grid_constant__ CUtensorMap mapZ16, FwdHArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Xc = smem;                 // split rows x[t]
  unsigned char* Xp = smem + TILE;          // split rows x[t-d]
  unsigned char* Zs = smem + 2 * TILE;      // split rows z (A operand of the dense product), later split rows of x'
  unsigned char* St = smem + 3 * TILE;      // fp32 staging: z (-> Zcat), later x'
  unsigned char* W0 = smem + 4 * TILE;      // [64][hi|lo] past tap (filter | gate)
  unsigned char* W1 = W0 + 8192;            // current tap
  unsigned char* Wd = W1 + 8192;            // [32][hi|lo] dense^T
  unsigned char* Zh = W0 + IMG_H;           // fp16 staging of z: [128 rows][32 halfs], 64B-swizzled
  __shared__ __align__(8) uint64_t bar_tma, bar_m1, bar_m2, bar_w;
  __shared__ uint32_t tmem_slot;
  __shared__ float pb_s[64];
  __shared__ float bd_s[32];

  // thread (r, half): time step t0 + r, channels [16*half, 16*half + 16)
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;
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
  if (tid == 0) {
    mbar_expect_tx(&bar_w, IMG_H);
    bulk_g2s(W0, a.img, IMG_H, &bar_w);
  }
  const int n_tt = (a.T + TM - 1) / TM;
  const int n_tiles = a.B * n_tt;
  auto issue_loads = [&](int tile) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    mbar_expect_tx(&bar_tma, 2 * TILE);
    tma_load_3d(Xc, &mapX, &bar_tma, 0, t0, b);
    tma_load_3d(Xp, &mapX, &bar_tma, 0, t0 - a.d, b);       // rows before the start of the window arrive as zeros
  };
  pdl_wait();      // the previous layer's output is complete and visible from here on
  if (a.pdl_next) pdl_trigger();
  if (tid == 0 && (int)blockIdx.x < n_tiles) issue_loads(blockIdx.x);
  mbar_wait(&bar_w, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16 * half;
  constexpr uint32_t ID64 = idesc_f16(128, 64), ID32 = idesc_f16(128, 32);
  const uint64_t dXc = kmajor_desc(smem_u32(Xc)), dXp = kmajor_desc(smem_u32(Xp)), dZs = kmajor_desc(smem_u32(Zs));
  const uint64_t dW0 = kmajor_desc(smem_u32(W0)), dW1 = kmajor_desc(smem_u32(W1)), dWd = kmajor_desc(smem_u32(Wd));
  // this thread's four 16-byte chunks of a split row: hi channels [16 half, +16) and the matching lo channels
  const uint32_t row_off = (uint32_t)r * 128;
  const uint32_t ch0 = (uint32_t)((2 * half) ^ (r & 7)) << 4, ch1 = (uint32_t)((2 * half + 1) ^ (r & 7)) << 4;
  const uint32_t cl0 = (uint32_t)((4 + 2 * half) ^ (r & 7)) << 4, cl1 = (uint32_t)((5 + 2 * half) ^ (r & 7)) << 4;

  int it = 0, pb_batch = -1;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = tile / n_tt, t0 = (tile - b * n_tt) * TM;
    const uint32_t par = it & 1;
    TLH(0);
    if (b != pb_batch) {   // block-uniform: conditioning + bias row of this batch element
      __syncthreads();
      if (tid < 64) pb_s[tid] = a.prebias[(size_t)b * 64 + tid];
      pb_batch = b;
      __syncthreads();      // (readers sit behind mbarrier waits only: without this a slow writer warp races them)
    }
    mbar_wait(&bar_tma, par);
    TLH(1);
    if (tid == 0) {
      tc_fence_after();
      mma_split(tmem, dXp, dW0, ID64, true);      // x[t-d] . W[0]
      mma_split(tmem, dXc, dW1, ID64, false);     // x[t]   . W[1]
      bulk_wait_read0();                          // the previous tile's output stores have left St / Zs
      mma_commit(&bar_m1);
    }
    // this thread's part of x[t] (for the residual sum), reassembled from its split row while the MMAs run
    float xo[16];
    {
      const uint4 h0 = *reinterpret_cast<const uint4*>(Xc + row_off + ch0), h1 = *reinterpret_cast<const uint4*>(Xc + row_off + ch1);
      const uint4 l0 = *reinterpret_cast<const uint4*>(Xc + row_off + cl0), l1 = *reinterpret_cast<const uint4*>(Xc + row_off + cl1);
      const __half2* hh0 = reinterpret_cast<const __half2*>(&h0);
      const __half2* hh1 = reinterpret_cast<const __half2*>(&h1);
      const __half2* ll0 = reinterpret_cast<const __half2*>(&l0);
      const __half2* ll1 = reinterpret_cast<const __half2*>(&l1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a0 = __half22float2(hh0[j]), b0 = __half22float2(ll0[j]);
        const float2 a1 = __half22float2(hh1[j]), b1 = __half22float2(ll1[j]);
        xo[2 * j] = a0.x + b0.x; xo[2 * j + 1] = a0.y + b0.y;
        xo[8 + 2 * j] = a1.x + b1.x; xo[8 + 2 * j + 1] = a1.y + b1.y;
      }
    }
    TLH(2);
    mbar_wait(&bar_m1, par);
    TLH(3);
    tc_fence_after();
    float z[16];
    {
      uint32_t fv[16], gv[16];
      tmem_ld16(lane_addr + 0, fv);
      tmem_ld16(lane_addr + 32, gv);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        z[j] = gated_fast(__uint_as_float(fv[j]) + pb_s[16 * half + j], __uint_as_float(gv[j]) + pb_s[32 + 16 * half + j]);
    }
    // z -> staging (tf32-rounded: it feeds the single-pass skip GEMM) and, split, the A operand of the dense product
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      *reinterpret_cast<float4*>(St + row_off + ((uint32_t)((4 * half + jj) ^ (r & 7)) << 4)) =
          make_float4(round_tf32(z[4 * jj]), round_tf32(z[4 * jj + 1]), round_tf32(z[4 * jj + 2]), round_tf32(z[4 * jj + 3]));
    if (a.z16) {
      __align__(16) __half zq[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) zq[j] = __float2half_rn(z[j]);
      unsigned char* zr = Zh + (uint32_t)r * 64;
      *reinterpret_cast<uint4*>(zr + ((uint32_t)((2 * half) ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(zq);
      *reinterpret_cast<uint4*>(zr + ((uint32_t)((2 * half + 1) ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(zq + 8);
    }
    if (!a.is_last) {
      __half zh[16], zl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split_h(z[j], zh[j], zl[j]);
      *reinterpret_cast<uint4*>(Zs + row_off + ch0) = *reinterpret_cast<const uint4*>(zh);
      *reinterpret_cast<uint4*>(Zs + row_off + ch1) = *reinterpret_cast<const uint4*>(zh + 8);
      *reinterpret_cast<uint4*>(Zs + row_off + cl0) = *reinterpret_cast<const uint4*>(zl);
      *reinterpret_cast<uint4*>(Zs + row_off + cl1) = *reinterpret_cast<const uint4*>(zl + 8);
    }
    fence_async_smem();
    tc_fence_before();
    TLH(4);
    __syncthreads();      // every thread has read its x row and written its z row
    TLH(5);
    if (tid == 0) {
      tma_store_3d(&mapZ, St, a.zcol, t0, b);      // rows past the end of the window are clipped by the tensor map
      if (a.z16) tma_store_3d(&mapZ16, Zh, a.zcol, t0, b);
      bulk_commit();
      if (tile + (int)gridDim.x < n_tiles) issue_loads(tile + gridDim.x);     // both input tiles are free
      if (!a.is_last) {
        tc_fence_after();
        mma_split(tmem + 64, dZs, dWd, ID32, true);      // z . Wd
        bulk_wait_read0();                               // the z store has left St before x' is staged there
        mma_commit(&bar_m2);
      }
    }
    if (!a.is_last) {
      mbar_wait(&bar_m2, par);
      TLH(6);
      tc_fence_after();
      uint32_t ov[16];
      tmem_ld16(lane_addr + 64, ov);
      float xn[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) xn[j] = xo[j] + __uint_as_float(ov[j]) + bd_s[16 * half + j];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        *reinterpret_cast<float4*>(St + row_off + ((uint32_t)((4 * half + jj) ^ (r & 7)) << 4)) =
            make_float4(xn[4 * jj], xn[4 * jj + 1], xn[4 * jj + 2], xn[4 * jj + 3]);
      __half xh[16], xl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split_h(xn[j], xh[j], xl[j]);
      *reinterpret_cast<uint4*>(Zs + row_off + ch0) = *reinterpret_cast<const uint4*>(xh);
      *reinterpret_cast<uint4*>(Zs + row_off + ch1) = *reinterpret_cast<const uint4*>(xh + 8);
      *reinterpret_cast<uint4*>(Zs + row_off + cl0) = *reinterpret_cast<const uint4*>(xl);
      *reinterpret_cast<uint4*>(Zs + row_off + cl1) = *reinterpret_cast<const uint4*>(xl + 8);
      fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();   // staging tiles complete; the TMEM columns are free for the next tile
    TLH(7);
    if (tid == 0 && !a.is_last) {
      tma_store_3d(&mapXo, St, 0, t0, b);      // fp32 x' (kept for the backward pass)
      tma_store_3d(&mapXs, Zs, 0, t0, b);      // split x' (next layer's operand rows)
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait0();      // every output store has landed before the grid counts as complete
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int block_fwd_h(const void* xs_in, void* xs_out, float* xout, float* zcat, void* zcat16, int ldz, int zcol, const unsigned char* img,
                const float* prebias, const float* dense_bias, int B, int T, int d, int is_last, int pdl_next,
                cudaStream_t st) {
  CUtensorMap mapX, mapZ, mapXo, mapXs, mapZ16;
  int rc = make_map_split(&mapX, (const __half*)xs_in, B, T);
  if (rc) return rc;
  rc = make_map_3d(&mapZ, zcat, B, T, ldz, ldz, TM);
  if (rc) return rc;
  mapXo = mapZ;
  mapXs = mapX;
  mapZ16 = mapX;
  if (zcat16) {
    rc = make_map_z16(&mapZ16, (const __half*)zcat16, B, T, ldz);
    if (rc) return rc;
  }
  if (!is_last) {
    rc = make_map_3d(&mapXo, xout, B, T, C, C, TM);
    if (rc) return rc;
    rc = make_map_split(&mapXs, (const __half*)xs_out, B, T);
    if (rc) return rc;
  }
  FwdHArgs a;
  a.img = img; a.prebias = prebias; a.dense_bias = dense_bias; a.B = B; a.T = T; a.d = d; a.is_last = is_last;
  a.zcol = zcol; a.pdl_next = pdl_next; a.z16 = zcat16 ? 1 : 0;
  a.timeline = (d == 32 && !is_last) ? g_timeline_h : nullptr;
  const size_t smem = 1024 + 4 * TILE + IMG_H + TM * 64;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(block_fwd_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int n_tiles = B * ((T + TM - 1) / TM);
  int grid = n_tiles;
  const int cap = 2 * sm_count();
  if (grid > cap) grid = cap;
  cudaError_t e = launch_pdl(block_fwd_h_kernel, dim3(grid), dim3(256), smem, st, mapX, mapZ, mapXo, mapXs, mapZ16, a);
  if (e != cudaSuccess) return (int)e;
  prof_mark(st, PT_BLOCK_FWD);
  return 0;
}

// ===================================================================================================================
// All forward layers in ONE persistent kernel ("chain").  The per-layer kernels above spend about half of each
// ~20 us launch outside their tile loop (grid ramp, weight fetch, store drain, inter-kernel dependency latency), 50
// times per step.  Here the CTAs stay resident for the whole stack and the layers are ordered by per-tile flags
// instead of kernel boundaries: tile (l, b, t0) needs the rows [t0-d, t0+128) of layer l-1, i.e. at most three tiles of
// it, and starts as soon as those have been published -- no grid-wide barrier anywhere.
//   warps 0-7  epilogue (thread = one time step x 16 channels, exactly as block_fwd_h_kernel)
//   warp  8    issuer   : claims work, waits for the producer tiles' flags, TMA loads, tcgen05 MMAs, weight images
//   warp  9    publisher: TMA stores of the staged outputs, then (stores complete) the tile's flag
// Work items are claimed from ONE global counter in (layer, tile) order.  That makes the kernel deadlock-free for any
// number of resident CTAs (no cooperative launch needed): an item only depends on items with a smaller index, those have
// been claimed by CTAs that are running, and a running CTA only ever waits on smaller items -- the issuer blocks on the
// flags of its next item only after it has issued everything its current one needs.  It also balances the 2.64 tiles
// per CTA per layer that the static launches round up to 3.
// The split rows travel through a ring of RING buffers [slot][B][T][hi|lo]; layer l reads slot l % RING and writes slot
// (l+1) % RING.  Before overwriting rows of a slot the publisher checks the flags of the layer-(l+1-RING) tiles that
// read them (write-after-read), which are RING-1 layers behind and practically always set.
constexpr int RING = 4;      // default ring depth (forward only); training keeps every layer's rows: ring = L + 1
constexpr int CH_THREADS = 320;

struct ChainArgs {
  const unsigned char* img;        // [L] weight images (IMG_H bytes each)
  const float* prebias;            // [L][B][64]
  const float* dense_bias;         // [L][32] or null
  unsigned int* flags;             // [L][n_tiles] tile flags, then the work counter; zeroed before the launch
  int L, B, T, n_tt;
  int has_xout, z16;
  int z32;                         // write z as fp32 into Zcat (off when every consumer reads the fp16 copy)
  int ring;                        // slots of the split-row ring (>= 2; L + 1: nothing is ever overwritten)
  int last_dense;                  // the last layer also computes its dense / residual output (stand-alone wn_block_fwd)
  long long* timeline;             // debug: cycles CTA 0 spent in each kind of wait (wn_debug_timeline)
  int dil[WN_MAX_LAYERS];
};

__global__ void __launch_bounds__(384, 2)      // 10 warps land 3+3 on one scheduler: 6 x 32 x regs <= 16384 needs <= 80 registers
block_fwd_chain_kernel(const __grid_constant__ CUtensorMap mapXS, const __grid_constant__ CUtensorMap mapZ,
                       const __grid_constant__ CUtensorMap mapXo, const __grid_constant__ CUtensorMap mapZ16,
                       const __grid_constant__ ChainArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Xc = smem;                 // split rows x[t]
  unsigned char* Xp = smem + TILE;          // split rows x[t-d]
  unsigned char* Zs = smem + 2 * TILE;      // split rows z (A operand of the dense product), later split rows of x'
  unsigned char* Sz = smem + 3 * TILE;      // fp32 staging of z (-> Zcat)
  unsigned char* Sx = smem + 4 * TILE;      // fp32 staging of x'
  unsigned char* W0 = smem + 5 * TILE;      // [64][hi|lo] past tap (filter | gate)
  unsigned char* W1 = W0 + 8192;            // current tap
  unsigned char* Wd = W1 + 8192;            // [32][hi|lo] dense^T
  unsigned char* Zh = W0 + IMG_H;           // fp16 staging of z: [128 rows][32 halfs], 64B-swizzled
  __shared__ __align__(8) uint64_t bar_tma, bar_m1, bar_m2, bar_w, bar_z, bar_o, bar_sfree, bar_xr;
  __shared__ uint32_t tmem_slot;
  __shared__ int item_s[4];                 // work item of tile i in item_s[i & 3] (-1: no more work), published through bar_tma
  __shared__ float pb_s[64];
  __shared__ float bd_s[32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_m1, 1);
    mbar_init(&bar_m2, 1);
    mbar_init(&bar_w, 1);
    mbar_init(&bar_z, 256);
    mbar_init(&bar_o, 257);      // 256 epilogue threads + the issuer (write-after-read check of the ring slot)
    mbar_init(&bar_sfree, 1);
    mbar_init(&bar_xr, 256);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_tiles = a.B * a.n_tt;
  const int n_items = a.L * n_tiles;

  if (warp == 8) {
    // ------------------------------------------------------------------------------------------------ issuer
    if (lane == 0) {
      constexpr uint32_t ID64 = idesc_f16(128, 64), ID32 = idesc_f16(128, 32);
      const uint64_t dXc = kmajor_desc(smem_u32(Xc)), dXp = kmajor_desc(smem_u32(Xp)), dZs = kmajor_desc(smem_u32(Zs));
      const uint64_t dW0 = kmajor_desc(smem_u32(W0)), dW1 = kmajor_desc(smem_u32(W1)), dWd = kmajor_desc(smem_u32(Wd));
      unsigned int* counter = a.flags + n_items;
      auto claim = [&]() -> int {
        const unsigned int w = atomicAdd(counter, 1u);
        return w < (unsigned int)n_items ? (int)w : -1;
      };
      const unsigned int *f0 = nullptr, *f1 = nullptr, *f2 = nullptr;      // flags of the producer tiles of the next item
      auto producer_flags = [&](int item) {
        const int l = item / n_tiles, j = item - l * n_tiles;
        const int b = j / a.n_tt, tt = j - b * a.n_tt, t0 = tt * TM, d = a.dil[l];
        if (l == 0) { f0 = nullptr; return; }
        const unsigned int* f = a.flags + (size_t)(l - 1) * n_tiles + (size_t)b * a.n_tt;
        f0 = f1 = f2 = f + tt;
        const int hi_t = t0 - d + TM - 1;
        if (hi_t >= 0) {
          const int lo_t = t0 - d > 0 ? t0 - d : 0;
          f1 = f + lo_t / TM;
          f2 = f + (hi_t < a.T ? hi_t : a.T - 1) / TM;
        }
      };
      auto issue_loads = [&](int item, bool ready) {     // producer tiles of layer l-1 published -> load both taps
        const int l = item / n_tiles, j = item - l * n_tiles;
        const int b = j / a.n_tt, tt = j - b * a.n_tt, t0 = tt * TM, d = a.dil[l];
        // (acquire, then TMA loads issued by this thread -- the pattern of griddepcontrol.wait + TMA; a fence.proxy.async
        // here was measured at ~2500 cycles on the critical path of every tile)
        if (f0 && !ready) wait_flags(f0, f1, f2);
        const int slot = l % a.ring;
        mbar_expect_tx(&bar_tma, 2 * TILE);      // (release: publishes item_s to the waiters of this phase)
        tma_load_3d(Xc, &mapXS, &bar_tma, 0, t0, slot * a.B + b);
        tma_load_3d(Xp, &mapXS, &bar_tma, 0, t0 - d, slot * a.B + b);     // rows before the window start arrive as zeros
      };
      long long c_flag = 0, c_w = 0, c_tma = 0, c_z = 0, c_m2 = 0, c_issue = 0, t_a, t_start = clock64();
      int item = claim();
      item_s[0] = item;
      uint32_t i = 0, wphase = 0;
      if (item < 0) {
        mbar_arrive(&bar_tma);
      } else {
        mbar_expect_tx(&bar_w, IMG_H);
        bulk_g2s(W0, a.img + (size_t)(item / n_tiles) * IMG_H, IMG_H, &bar_w);
        producer_flags(item);
        issue_loads(item, false);
        mbar_wait(&bar_w, 0);
        mbar_wait(&bar_tma, 0);
        tc_fence_after();
        mma_split(tmem, dXp, dW0, ID64, true);      // x[t-d] . W[0]
        mma_split(tmem, dXc, dW1, ID64, false);     // x[t]   . W[1]
        mma_commit(&bar_m1);
      }
      // Invariant at the loop top: the first product of tile i has been issued.
      while (item >= 0) {
        const uint32_t par = i & 1;
        const int l = item / n_tiles;
        const bool last = (l == a.L - 1) && !a.last_dense;
        // the next item is claimed and its producers' flags are looked at while the epilogue warps work on this one;
        // BLOCKING on flags has to wait until everything this tile needs has been issued (see the header)
        const int nx = claim();
        item_s[(i + 1) & 3] = nx;
        const int nl = nx >= 0 ? nx / n_tiles : l;
        bool ready = true;
        if (nx >= 0) {
          producer_flags(nx);
          ready = (f0 == nullptr);
        }
        // write-after-read: the ring slot this tile's x' goes to was read by layer l + 1 - RING (own rows and the
        // shifted tap of the tiles d later); those tiles are RING - 1 layers behind and practically always published
        const unsigned int *w0 = nullptr, *w1 = nullptr, *w2 = nullptr;
        if (!last && l + 1 - a.ring >= 0) {
          const int lr = l + 1 - a.ring, j = item - l * n_tiles;
          const int b = j / a.n_tt, tt = j - b * a.n_tt, t0 = tt * TM, dr = a.dil[lr];
          const unsigned int* f = a.flags + (size_t)lr * n_tiles + (size_t)b * a.n_tt;
          const int ta = (t0 + dr) / TM, tb = (t0 + TM - 1 + dr) / TM;
          w0 = f + tt; w1 = f + (ta < a.n_tt ? ta : tt); w2 = f + (tb < a.n_tt ? tb : tt);
        }
        bool war_ok = (w0 == nullptr);
        // both input tiles are free as soon as the first product has read them and every epilogue thread has its x row:
        // the next tile's loads (and W0/W1 of a new layer) start here, a whole epilogue ahead of the z barrier
        t_a = clock64();
        mbar_wait(&bar_m1, par);
        mbar_wait(&bar_xr, par);
        c_tma += clock64() - t_a;
        bool loaded = (nx < 0);
        auto early = [&]() {
          if (nl != l) {
            mbar_expect_tx(&bar_w, IMG_H);
            bulk_g2s(W0, a.img + (size_t)nl * IMG_H, 16384, &bar_w);
          }
          issue_loads(nx, true);
          loaded = true;
        };
        t_a = clock64();
        for (;;) {
          if (!ready) ready = flags_set(f0, f1, f2);
          if (ready && !loaded) early();
          if (!war_ok) war_ok = flags_set(w0, w1, w2);
          if ((loaded && war_ok) || mbar_test(&bar_z, par)) break;
        }
        mbar_wait(&bar_z, par);                     // z staged in Zs
        c_z += clock64() - t_a;
        t_a = clock64();
        tc_fence_after();
        if (!last) {
          mma_split(tmem + 64, dZs, dWd, ID32, true);      // z . Wd
          mma_commit(&bar_m2);
          if (!war_ok) wait_flags(w0, w1, w2);
          mbar_arrive(&bar_o);                              // the publisher may overwrite the slot rows
        }
        c_issue += clock64() - t_a;
        if (nx < 0) {
          mbar_arrive(&bar_tma);      // completes the next phase without data: the other warps see "no more work"
          break;
        }
        t_a = clock64();
        if (!loaded) {
          wait_flags(f0, f1, f2);
          early();
        }
        c_flag += clock64() - t_a;
        if (nl != l) {      // Wd is free once the dense product has finished
          t_a = clock64();
          if (!last) mbar_wait(&bar_m2, par);
          c_m2 += clock64() - t_a;
          bulk_g2s(Wd, a.img + (size_t)nl * IMG_H + 16384, 4096, &bar_w);
          wphase ^= 1;
          t_a = clock64();
          mbar_wait(&bar_w, wphase);
          c_w += clock64() - t_a;
        }
        t_a = clock64();
        mbar_wait(&bar_tma, par ^ 1u);
        c_tma += clock64() - t_a;
        tc_fence_after();
        mma_split(tmem, dXp, dW0, ID64, true);      // (the accumulator columns are free: every thread has staged its z)
        mma_split(tmem, dXc, dW1, ID64, false);
        mma_commit(&bar_m1);
        item = nx;
        ++i;
      }
      if (a.timeline && blockIdx.x == 0) {
        a.timeline[0] = clock64() - t_start; a.timeline[1] = c_flag; a.timeline[2] = c_w; a.timeline[3] = c_tma;
        a.timeline[4] = c_z; a.timeline[5] = c_m2; a.timeline[6] = c_issue; a.timeline[7] = i; a.timeline[8] = gridDim.x;
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------------------------------------ publisher
    if (lane == 0) {
      long long p_z = 0, p_war = 0, p_o = 0, p_read = 0, p_done = 0, t_a;
      const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
      int pending = -1;      // tile whose flag waits for its stores (published after the next tile's z store)
      for (uint32_t i = 0;; ++i) {
        const uint32_t par = i & 1;
        // (not bar_tma: its phase i+1 can complete while this thread is still publishing tile i-1, and a parity wait
        // that is two phases late never returns.  bar_z cannot run ahead: phase i+1 needs this thread's bar_sfree(i).)
        t_a = clock64();
        mbar_wait(&bar_z, par);
        p_z += clock64() - t_a;
        const int item = item_s[i & 3];
        if (item < 0) break;
        const int l = item / n_tiles, j = item - l * n_tiles;
        const bool last = (l == a.L - 1) && !a.last_dense;
        const int b = j / a.n_tt, tt = j - b * a.n_tt, t0 = tt * TM;
        if (a.z32) tma_store_3d_hint(&mapZ, Sz, l * C, t0, b, pol_stream);      // rows past the end of the window are clipped by the tensor map
        if (a.z16) tma_store_3d_hint(&mapZ16, Zh, l * C, t0, b, pol_stream);
        bulk_commit();
        if (pending >= 0) {      // the previous tile's stores: every group but the one just committed has completed
          t_a = clock64();
          asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
          __threadfence();
          st_release(a.flags + pending, 1u);
          pending = -1;
          p_war += clock64() - t_a;
        }
        if (!last) {
          t_a = clock64();
          mbar_wait(&bar_o, par);                  // x' staged, and the issuer has seen the old readers of the slot finish
          p_o += clock64() - t_a;
          if (a.has_xout) tma_store_3d_hint(&mapXo, Sx, 0, t0, (l + 1) * a.B + b, pol_stream);      // fp32 x' (kept for the backward pass)
          tma_store_3d_hint(&mapXS, Zs, 0, t0, ((l + 1) % a.ring) * a.B + b, pol_keep);             // split x' (next layer's operand rows)
          bulk_commit();
        }
        t_a = clock64();
        bulk_wait_read0();
        p_read += clock64() - t_a;
        mbar_arrive(&bar_sfree);
        if (!last) {
          // publish the tile once its stores have completed.  When the next tile is already staged its z store goes
          // first and the flag follows it (wait_group 1 above); only then: a flag held back behind a tile that is not
          // staged yet could be the very flag that tile's loads wait for.
          if (mbar_test(&bar_z, par ^ 1u)) {
            pending = item;
          } else {
            t_a = clock64();
            bulk_wait0();                 // the stores have completed and are visible to this thread ...
            __threadfence();
            st_release(a.flags + item, 1u);      // ... publish the tile
            p_done += clock64() - t_a;
          }
        }
      }
      bulk_wait0();
      if (pending >= 0) {
        __threadfence();
        st_release(a.flags + pending, 1u);
      }
      if (a.timeline && blockIdx.x == 0) {
        a.timeline[10] = p_z; a.timeline[11] = p_war; a.timeline[12] = p_o; a.timeline[13] = p_read; a.timeline[14] = p_done;
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
      mbar_wait(&bar_tma, par);
      const int item = item_s[i & 3];
      if (item < 0) {
        // hands the end marker on to the publisher -- once it has consumed the previous phase of bar_z (its
        // bar_sfree arrival follows its bar_z wait): a parity wait that falls two phases behind never returns
        if (i > 0) mbar_wait(&bar_sfree, (i - 1) & 1);
        mbar_arrive(&bar_z);
        break;
      }
      const int l = item / n_tiles, j = item - l * n_tiles;
      const bool last = (l == a.L - 1) && !a.last_dense;
      const int b = j / a.n_tt;
      if (l * a.B + b != key) {      // uniform over the epilogue warps: bias rows of this (layer, batch element)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < 64) pb_s[tid] = a.prebias[((size_t)l * a.B + b) * 64 + tid];
        else if (tid < 96) bd_s[tid - 64] = a.dense_bias ? a.dense_bias[(size_t)l * C + tid - 64] : 0.f;
        key = l * a.B + b;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      float xo[16];
      {
        const uint4 h0 = *reinterpret_cast<const uint4*>(Xc + row_off + ch0), h1 = *reinterpret_cast<const uint4*>(Xc + row_off + ch1);
        const uint4 l0 = *reinterpret_cast<const uint4*>(Xc + row_off + cl0), l1 = *reinterpret_cast<const uint4*>(Xc + row_off + cl1);
        const __half2* hh0 = reinterpret_cast<const __half2*>(&h0);
        const __half2* hh1 = reinterpret_cast<const __half2*>(&h1);
        const __half2* ll0 = reinterpret_cast<const __half2*>(&l0);
        const __half2* ll1 = reinterpret_cast<const __half2*>(&l1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a0 = __half22float2(hh0[q]), b0 = __half22float2(ll0[q]);
          const float2 a1 = __half22float2(hh1[q]), b1 = __half22float2(ll1[q]);
          xo[2 * q] = a0.x + b0.x; xo[2 * q + 1] = a0.y + b0.y;
          xo[8 + 2 * q] = a1.x + b1.x; xo[8 + 2 * q + 1] = a1.y + b1.y;
        }
      }
      mbar_arrive(&bar_xr);      // this thread no longer needs the x tile
      mbar_wait(&bar_m1, par);
      tc_fence_after();
      float z[16];
      {
        uint32_t fv[16], gv[16];
        tmem_ld16(lane_addr + 0, fv);
        tmem_ld16(lane_addr + 32, gv);
#pragma unroll
        for (int q = 0; q < 16; ++q)
          z[q] = gated_fast(__uint_as_float(fv[q]) + pb_s[16 * half + q], __uint_as_float(gv[q]) + pb_s[32 + 16 * half + q]);
      }
      if (i > 0) mbar_wait(&bar_sfree, (i - 1) & 1);      // the previous tile's stores have left the staging tiles
      if (a.z32) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          *reinterpret_cast<float4*>(Sz + row_off + ((uint32_t)((4 * half + jj) ^ (r & 7)) << 4)) =
              make_float4(round_tf32(z[4 * jj]), round_tf32(z[4 * jj + 1]), round_tf32(z[4 * jj + 2]), round_tf32(z[4 * jj + 3]));
      }
      if (a.z16) {
        __align__(16) __half zq[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) zq[q] = __float2half_rn(z[q]);
        unsigned char* zr = Zh + (uint32_t)r * 64;
        *reinterpret_cast<uint4*>(zr + ((uint32_t)((2 * half) ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(zq);
        *reinterpret_cast<uint4*>(zr + ((uint32_t)((2 * half + 1) ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(zq + 8);
      }
      if (!last) {
        __align__(16) __half zh[16], zl[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) split_h(z[q], zh[q], zl[q]);
        *reinterpret_cast<uint4*>(Zs + row_off + ch0) = *reinterpret_cast<const uint4*>(zh);
        *reinterpret_cast<uint4*>(Zs + row_off + ch1) = *reinterpret_cast<const uint4*>(zh + 8);
        *reinterpret_cast<uint4*>(Zs + row_off + cl0) = *reinterpret_cast<const uint4*>(zl);
        *reinterpret_cast<uint4*>(Zs + row_off + cl1) = *reinterpret_cast<const uint4*>(zl + 8);
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(&bar_z);
      if (!last) {
        mbar_wait(&bar_m2, par);
        tc_fence_after();
        uint32_t ov[16];
        tmem_ld16(lane_addr + 64, ov);
        float xn[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) xn[q] = xo[q] + __uint_as_float(ov[q]) + bd_s[16 * half + q];
        if (a.has_xout) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            *reinterpret_cast<float4*>(Sx + row_off + ((uint32_t)((4 * half + jj) ^ (r & 7)) << 4)) =
                make_float4(xn[4 * jj], xn[4 * jj + 1], xn[4 * jj + 2], xn[4 * jj + 3]);
        }
        __align__(16) __half xh[16], xl[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) split_h(xn[q], xh[q], xl[q]);
        *reinterpret_cast<uint4*>(Zs + row_off + ch0) = *reinterpret_cast<const uint4*>(xh);
        *reinterpret_cast<uint4*>(Zs + row_off + ch1) = *reinterpret_cast<const uint4*>(xh + 8);
        *reinterpret_cast<uint4*>(Zs + row_off + cl0) = *reinterpret_cast<const uint4*>(xl);
        *reinterpret_cast<uint4*>(Zs + row_off + cl1) = *reinterpret_cast<const uint4*>(xl + 8);
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(&bar_o);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 128);
}

// xs_ring: RING x [B][T][hi 32 | lo 32] fp16 (slot 0 holds the input of layer 0); xall: fp32 [L][B][T][32] with
// xall[0] the input of layer 0 (layer l writes xall[l+1]; null: forward only); flags: L * B * ceil(T/128) + 1 words.
int block_fwd_chain(void* xs_ring, float* xall, float* zcat, void* zcat16, int ldz, const unsigned char* img,
                    const float* prebias, const float* dense_bias, const int* dilations, int L, int B, int T,
                    unsigned int* flags, cudaStream_t st, int ring, int last_dense, int zcols) {
  if (L < 1 || L > WN_MAX_LAYERS) return -1;
  if (ring <= 0) ring = RING;
  if (ring < 2) return -1;
  if (zcols <= 0) zcols = ldz;
  CUtensorMap mapXS, mapZ, mapXo, mapZ16;
  int rc = make_map_split(&mapXS, (const __half*)xs_ring, (int64_t)ring * B, T);
  if (rc) return rc;
  if (!zcat && !zcat16) return -1;
  mapZ = mapXS;
  if (zcat) {
    rc = make_map_3d(&mapZ, zcat, B, T, zcols, ldz, TM);
    if (rc) return rc;
  }
  mapXo = mapXS;
  if (xall) {
    rc = make_map_3d(&mapXo, xall, (int64_t)(L + (last_dense ? 1 : 0)) * B, T, C, C, TM);
    if (rc) return rc;
  }
  mapZ16 = mapXS;
  if (zcat16) {
    rc = make_map_z16(&mapZ16, (const __half*)zcat16, B, T, ldz);
    if (rc) return rc;
  }
  ChainArgs a;
  a.img = img; a.prebias = prebias; a.dense_bias = dense_bias; a.flags = flags;
  a.L = L; a.B = B; a.T = T; a.n_tt = (T + TM - 1) / TM;
  a.has_xout = xall ? 1 : 0; a.z16 = zcat16 ? 1 : 0;
  a.ring = ring; a.last_dense = last_dense ? 1 : 0; a.z32 = zcat ? 1 : 0;
  a.timeline = g_timeline_h;
  for (int l = 0; l < WN_MAX_LAYERS; ++l) a.dil[l] = l < L ? dilations[l] : 0;
  const size_t smem = 1024 + 5 * TILE + IMG_H + TM * 64;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(block_fwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -5;
    cudaFuncSetAttribute(block_fwd_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  const int64_t n_items = (int64_t)L * B * a.n_tt;
  if (n_items >= (1ll << 31) - 4096) return -1;
  int grid = 2 * sm_count();      // two CTAs fit an SM (111 KB shared memory, 80 registers x 320 threads each)
  if (grid > B * a.n_tt) grid = B * a.n_tt;
  cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)(n_items + 1) * sizeof(unsigned int), st);
  if (e != cudaSuccess) return (int)e;
  block_fwd_chain_kernel<<<grid, CH_THREADS, smem, st>>>(mapXS, mapZ, mapXo, mapZ16, a);
  WN_CHECK_LAUNCH();
  prof_mark(st, PT_BLOCK_FWD);
  return 0;
}
int64_t block_fwd_chain_ring_bytes(int64_t M, int ring) { return (int64_t)(ring > 0 ? ring : RING) * M * 128; }

int block_fwd_h_set_trap_info(unsigned int* p) { return umma::set_trap_info_tu(p); }

}  // namespace wn
