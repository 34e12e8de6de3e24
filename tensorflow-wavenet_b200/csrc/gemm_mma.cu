// TF32 tensor-core GEMM (mma.sync m16n8k8, fp32 accumulate) with the epilogues the skip /
// post-processing path needs (reference call sites: wavenet/model.py:304-305,430-440 and
// the TF autodiff of them).  Three operand forms, all row-major storage:
//   NN: C[M,N] = A[M,K]   . B[K,N]        forward 1x1 convolutions (skip-sum GEMM, post1, post2)
//   NT: C[M,N] = A[M,K]   . B[N,K]^T      input gradients
//   TN: C[M,N] = A[Kr,M]^T . B[Kr,N]      weight gradients (Kr = B*T rows, split over grid.z,
//                                          accumulated with atomics into a zeroed C)
// 128x128x32 CTA tile, 8 warps (2x4, 64x32 each), 3-stage cp.async pipeline.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace wn {

namespace {
constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int PK = BK + 4;     // pitch of a K-contiguous tile  [128][36]
constexpr int PMN = BM + 8;    // pitch of an MN-contiguous tile [32][136]
constexpr int TILE_FLOATS = 128 * PK;  // 4608 >= 32*136 = 4352
constexpr int STAGE_FLOATS = 2 * TILE_FLOATS;

// K-contiguous tile: rows = M or N index (128), cols = k (32)
__device__ __forceinline__ void load_kmajor(float* s, const float* __restrict__ gptr, int ld, int row0,
                                            int row_lim, int k0, int k_lim) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = threadIdx.x + it * 256;
    const int r = i >> 3, kc = (i & 7) * 4;
    const bool v = (row0 + r < row_lim) && (k0 + kc < k_lim);
    const float* src = v ? gptr + (size_t)(row0 + r) * ld + k0 + kc : gptr;
    cp_async16(s + r * PK + kc, src, v);
  }
}
// MN-contiguous tile: rows = k (32), cols = M or N index (128)
__device__ __forceinline__ void load_mnmajor(float* s, const float* __restrict__ gptr, int ld, int k0,
                                             int k_lim, int col0, int col_lim) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = threadIdx.x + it * 256;
    const int kr = i >> 5, c = (i & 31) * 4;
    const bool v = (k0 + kr < k_lim) && (col0 + c < col_lim);
    const float* src = v ? gptr + (size_t)(k0 + kr) * ld + col0 + c : gptr;
    cp_async16(s + kr * PMN + c, src, v);
  }
}
}  // namespace

template <int MODE>   // 0 NN, 1 NT, 2 TN
__global__ void __launch_bounds__(256, 2) gemm_tf32_kernel(GemmParams p) {
  extern __shared__ __align__(16) float smem[];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;

  // K range of this CTA (split-K only used by TN)
  const int nkb_total = (p.K + BK - 1) / BK;
  const int per = (nkb_total + gridDim.z - 1) / gridDim.z;
  const int kb_begin = blockIdx.z * per;
  int kb_end = kb_begin + per;
  if (kb_end > nkb_total) kb_end = nkb_total;
  const int nkb = kb_end - kb_begin;
  if (nkb <= 0) return;

  auto load_stage = [&](int stage, int kb) {
    float* sa = smem + stage * STAGE_FLOATS;
    float* sb = sa + TILE_FLOATS;
    const int k0 = kb * BK;
    if (MODE == 2) load_mnmajor(sa, p.A, p.lda, k0, p.K, m0, p.M);
    else load_kmajor(sa, p.A, p.lda, m0, p.M, k0, p.K);
    if (MODE == 1) load_kmajor(sb, p.B, p.ldb, n0, p.N, k0, p.K);
    else load_mnmajor(sb, p.B, p.ldb, k0, p.K, n0, p.N);
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s, kb_begin + s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kb + STAGES - 1;
      if (nxt < nkb) load_stage(nxt % STAGES, kb_begin + nxt);
      cp_async_commit();
    }
    const float* sa = smem + (kb % STAGES) * STAGE_FLOATS;
    const float* sb = sa + TILE_FLOATS;
#pragma unroll
    for (int ks = 0; ks < BK / 8; ++ks) {
      uint32_t af[4][4], bf[4][2];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (MODE == 2) {
          const float* a = sa + (8 * ks + t) * PMN + wm * 64 + 16 * mt + g;
          af[mt][0] = f2tf32(a[0]);
          af[mt][1] = f2tf32(a[8]);
          af[mt][2] = f2tf32(a[4 * PMN]);
          af[mt][3] = f2tf32(a[4 * PMN + 8]);
        } else {
          const float* a = sa + (wm * 64 + 16 * mt + g) * PK + 8 * ks + t;
          af[mt][0] = f2tf32(a[0]);
          af[mt][1] = f2tf32(a[8 * PK]);
          af[mt][2] = f2tf32(a[4]);
          af[mt][3] = f2tf32(a[8 * PK + 4]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (MODE == 1) {
          const float* b = sb + (wn * 32 + 8 * nt + g) * PK + 8 * ks + t;
          bf[nt][0] = f2tf32(b[0]);
          bf[nt][1] = f2tf32(b[4]);
        } else {
          const float* b = sb + (8 * ks + t) * PMN + wn * 32 + 8 * nt + g;
          bf[nt][0] = f2tf32(b[0]);
          bf[nt][1] = f2tf32(b[4 * PMN]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_tf32(acc[mt][nt], af[mt][0], af[mt][1], af[mt][2], af[mt][3], bf[nt][0], bf[nt][1]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ----
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = m0 + wm * 64 + 16 * mt + g + 8 * h;
      if (row >= p.M) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = n0 + wn * 32 + 8 * nt + 2 * t;
        if (col >= p.N) continue;
        float v0 = acc[mt][nt][2 * h], v1 = acc[mt][nt][2 * h + 1];
        if (p.flags & GEMM_ATOMIC) {
          atomicAdd(p.C + (size_t)row * p.ldc + col, v0);
          atomicAdd(p.C + (size_t)row * p.ldc + col + 1, v1);
          continue;
        }
        if (p.bias) {
          v0 += __ldg(p.bias + col);
          v1 += __ldg(p.bias + col + 1);
        }
        if (p.C2) *reinterpret_cast<float2*>(p.C2 + (size_t)row * p.ldc2 + col) = make_float2(v0, v1);
        if (p.flags & GEMM_RELU) {
          v0 = fmaxf(v0, 0.f);
          v1 = fmaxf(v1, 0.f);
        }
        if (p.aux) {   // gradient of relu: keep where the saved activation is positive
          float2 m = __ldg(reinterpret_cast<const float2*>(p.aux + (size_t)row * p.ldaux + col));
          v0 = m.x > 0.f ? v0 : 0.f;
          v1 = m.y > 0.f ? v1 : 0.f;
        }
        if (p.flags & GEMM_ROUND) {
          v0 = round_tf32(v0);
          v1 = round_tf32(v1);
        }
        *reinterpret_cast<float2*>(p.C + (size_t)row * p.ldc + col) = make_float2(v0, v1);
      }
    }
  }
}

int gemm_tf32(int mode, const GemmParams& p, int split_k, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return -1;
  if ((p.lda & 3) || (p.ldb & 3) || (p.N & 1) || (p.ldc & 1)) return -3;
  if (mode != 2 && (p.K & 3)) return -3;
  if (mode == 2 && (p.M & 3)) return -3;
  if (mode != 1 && (p.N & 3)) return -3;
  if (((uintptr_t)p.A & 15) || ((uintptr_t)p.B & 15) || ((uintptr_t)p.C & 7)) return -4;
  const size_t smem = sizeof(float) * STAGES * STAGE_FLOATS;
  static bool attr[3] = {false, false, false};
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, (p.flags & GEMM_ATOMIC) ? (split_k > 0 ? split_k : 1) : 1);
  if (mode == 0) {
    if (!attr[0]) { cudaFuncSetAttribute(gemm_tf32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr[0] = true; }
    gemm_tf32_kernel<0><<<grid, 256, smem, st>>>(p);
  } else if (mode == 1) {
    if (!attr[1]) { cudaFuncSetAttribute(gemm_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr[1] = true; }
    gemm_tf32_kernel<1><<<grid, 256, smem, st>>>(p);
  } else if (mode == 2) {
    if (!attr[2]) { cudaFuncSetAttribute(gemm_tf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr[2] = true; }
    gemm_tf32_kernel<2><<<grid, 256, smem, st>>>(p);
  } else {
    return -1;
  }
  WN_CHECK_LAUNCH();
  return 0;
}

// out[n] += sum_m A[m][n]   (bias gradients); out must be zeroed by the caller
// out[n] += sum_m A[m][n].  Block = 32 x 8 threads: a warp reads 512 contiguous bytes of one row (float4 per
// lane), the 8 warps take interleaved rows with 4 loads in flight each, then reduce through shared memory.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int lda, int M, int N,
                                                     float* __restrict__ out, int rows_per_cta) {
  __shared__ float4 red[8][32];
  const int col = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int r0 = blockIdx.y * rows_per_cta;
  int r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float4 s[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(r + 8 * u) * lda + col));
        s[u].x += v.x; s[u].y += v.y; s[u].z += v.z; s[u].w += v.w;
      }
    }
    for (; r < r1; r += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(A + (size_t)r * lda + col));
      s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
    }
  }
  red[threadIdx.y][threadIdx.x] = make_float4((s[0].x + s[1].x) + (s[2].x + s[3].x), (s[0].y + s[1].y) + (s[2].y + s[3].y),
                                              (s[0].z + s[1].z) + (s[2].z + s[3].z), (s[0].w + s[1].w) + (s[2].w + s[3].w));
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 v = red[w][threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    atomicAdd(out + col, t.x);
    atomicAdd(out + col + 1, t.y);
    atomicAdd(out + col + 2, t.z);
    atomicAdd(out + col + 3, t.w);
  }
}

// the same for an fp16 matrix in a scaled domain: out[n] += scale * sum_m A16[m][n]   (a lane owns 8 columns)
// partials != null: every row chunk writes its sums to partials[chunk][N] and colsum_finish_kernel adds them up.  Atomics
// on ONE address serialise at ~0.25 us each: 148 chunks adding to every column bounded the kernel at ~40 us for 100 MB.
__global__ void __launch_bounds__(256) colsum16_kernel(const __half* __restrict__ A, int lda, int M, int N, float scale,
                                                       float* __restrict__ out, int rows_per_cta, float* __restrict__ partials) {
  __shared__ float red[8][32][8];
  const int col = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  int r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float s[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) s[u] = 0.f;
  if (col < N) {
    auto add = [&](const uint4& v) {
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = __half22float2(h[u]);
        s[2 * u] += f.x; s[2 * u + 1] += f.y;
      }
    };
    int r = r0 + threadIdx.y;
    for (; r + 56 < r1; r += 64) {      // eight 16-byte loads in flight per thread (~64 KB per SM: enough to cover HBM latency)
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(A + (size_t)(r + 8 * u) * lda + col));
#pragma unroll
      for (int u = 0; u < 8; ++u) add(v[u]);
    }
    for (; r < r1; r += 8) add(__ldg(reinterpret_cast<const uint4*>(A + (size_t)r * lda + col)));
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) red[threadIdx.y][threadIdx.x][u] = s[u];
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x][u];
      if (partials) partials[(size_t)blockIdx.y * N + col + u] = t * scale;
      else atomicAdd(out + col + u, t * scale);
    }
  }
}
__global__ void colsum_finish_kernel(const float* __restrict__ partials, int chunks, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 3 < chunks; c += 4) {
    s0 += partials[(size_t)c * N + n]; s1 += partials[(size_t)(c + 1) * N + n];
    s2 += partials[(size_t)(c + 2) * N + n]; s3 += partials[(size_t)(c + 3) * N + n];
  }
  for (; c < chunks; ++c) s0 += partials[(size_t)c * N + n];
  out[n] += (s0 + s1) + (s2 + s3);
}
int64_t colsum16_scratch_floats(int N) { return (int64_t)(2 * sm_count() + 2) * N; }
int colsum16(const void* A16, int lda, int M, int N, float scale, float* out, float* scratch, cudaStream_t st) {
  if (M <= 0 || N <= 0) return -1;
  if ((N & 7) || (lda & 7) || ((uintptr_t)A16 & 15)) return -3;
  const int bx = (N + 255) / 256;
  // (every CTA ends with one atomicAdd per column: 1184 row chunks of a 256-column matrix serialised ~1200 atomics on
  // every address and the kernel took 78 us for 50 MB; two CTAs per SM keep enough loads in flight)
  // (more, smaller chunks were measured slower: the finishing kernel then walks 1184 partial rows per column)
  int chunks = (2 * sm_count() + bx - 1) / bx;
  int rows = (M + chunks - 1) / chunks;
  if (rows < 64) rows = 64;
  chunks = (M + rows - 1) / rows;
  if (chunks > 2 * sm_count() + 2) scratch = nullptr;
  colsum16_kernel<<<dim3(bx, chunks), dim3(32, 8), 0, st>>>((const __half*)A16, lda, M, N, scale, out, rows, scratch);
  WN_CHECK_LAUNCH();
  if (scratch) {
    colsum_finish_kernel<<<(N + 127) / 128, 128, 0, st>>>(scratch, chunks, N, out);
    WN_CHECK_LAUNCH();
  }
  return 0;
}

int colsum(const float* A, int lda, int M, int N, float* out, cudaStream_t st) {
  if (M <= 0 || N <= 0) return -1;
  if ((N & 3) || (lda & 3) || ((uintptr_t)A & 15)) return -3;
  const int bx = (N + 127) / 128;
  int chunks = (8 * sm_count() + bx - 1) / bx;
  int rows = (M + chunks - 1) / chunks;
  if (rows < 64) rows = 64;
  chunks = (M + rows - 1) / rows;
  colsum_kernel<<<dim3(bx, chunks), dim3(32, 8), 0, st>>>(A, lda, M, N, out, rows);
  WN_CHECK_LAUNCH();
  return 0;
}

}  // namespace wn
