// tcgen05 (5th-generation tensor core) TF32 GEMM for the skip / post-processing path, fp32 row-major
// storage, three operand forms (the same three the TF graph and its autodiff need):
//     NN: C[M,N]   = A[M,K] . B[K,N]          forward 1x1 convolutions      (A K-major, B MN-major)
//     NT: C[M,N]   = A[M,K] . B[N,K]^T        input gradients               (A, B K-major)
//     TN: C[M,N]  += A[Kr,M]^T . B[Kr,N]      weight gradients, Kr = B*T    (A, B MN-major, split-K + red.add)
// Reference call sites: wavenet/model.py:304-305,430-440 (1x1 convolutions = matmuls) and the
// TF autodiff of them (train.py:252).  Every operand is read as it lies in HBM: K-major tiles use
// TMA's 128-byte swizzle, MN-major tiles the 32-byte-atom flavour (the only one kind::tf32 accepts,
// probe/umma_probe.cu) -- no transposed copies of activations or weights exist.
//
// Structure (one 128 x BN output tile per CTA, BN in {128, 256}):
//   warp 0 : TMA producer   -- cp.async.bulk.tensor 2D boxes [128|BN rows x 32 floats] into a
//            4-stage ring of 128B-swizzled tiles, completion on "full" mbarriers
//   warp 1 : MMA issuer     -- one lane issues 4 x tcgen05.mma.kind::tf32 (K=8 each) per stage into a
//            128 x BN fp32 accumulator in TMEM; tcgen05.commit releases the stage ("empty" mbarrier)
//   warps 2-5 : epilogue    -- tcgen05.ld the accumulator (each warp its TMEM lane quadrant), apply
//            bias / relu / relu-gradient mask / tf32 rounding, store row-major and (optionally) a
//            transposed copy (coalesced: lanes hold consecutive rows), or red.add for split-K.
// Descriptor conventions were validated on B200 by probe/umma_probe.cu.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "umma_common.cuh"

namespace wn {

namespace {

constexpr int UM = 128;        // UMMA M (rows of A per tile)
constexpr int UK = 32;         // floats per stage along K (= one 128-byte swizzle row)
constexpr int STAGES = 4;
constexpr int THREADS = 192;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

// fp32 row-major [rows][cols] (row pitch ld floats); box = [box_rows][32 floats]; 128B swizzle; OOB -> 0
int make_map(CUtensorMap* m, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must fail loudly (trap -> launch error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 24); ++i) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, 128B swizzle: 8-row groups 1024 B apart (SBO), LBO unused (encoded 1), version 1
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

}  // namespace

struct UmmaParams {
  float* C; int ldc;              // row-major output (nullable when only CT is wanted)
  float* CT; int ldct;            // transposed copy CT[n][m] (nullable)
  float* C2; int ldc2;            // raw (acc + bias) second output (nullable)
  const float* bias;              // per column n (nullable)
  const float* aux; int ldaux;    // relu-gradient mask source (nullable)
  int M, N, K;
  int flags;                      // GEMM_RELU | GEMM_ROUND | GEMM_ATOMIC
  int k_per_split;                // K range per blockIdx.z (multiple of UK)
  int a_mn, b_mn;                 // operand majors: 0 = K contiguous, 1 = M / N contiguous
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, UmmaParams p) {
  constexpr uint32_t A_BYTES = UM * UK * 4;       // 16 KB
  constexpr uint32_t B_BYTES = BN * UK * 4;       // 16 / 32 KB
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * UM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * p.k_per_split;
  int k_end = k_begin + p.k_per_split;
  if (k_end > p.K) k_end = p.K;
  const int nk = (k_end - k_begin + UK - 1) / UK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0 && nk > 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        unsigned char* sa = smem + s * STAGE_BYTES;
        const int kc = k_begin + i * UK;
        if (p.a_mn) umma::tma_load_3d(sa, &mapA, &full_bar[s], 0, kc, m0 >> 5);     // [UM/32 blocks][32 k][32 m]
        else tma_load_2d(sa, &mapA, &full_bar[s], kc, m0);
        if (p.b_mn) umma::tma_load_3d(sa + A_BYTES, &mapB, &full_bar[s], 0, kc, n0 >> 5);
        else tma_load_2d(sa + A_BYTES, &mapB, &full_bar[s], kc, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nk > 0) {
      const uint32_t idesc = umma::idesc_tf32(UM, BN, p.a_mn, p.b_mn);
      const uint32_t astep = p.a_mn ? 64 : 2, bstep = p.b_mn ? 64 : 2;   // descriptor start-address step per K=8
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t da = p.a_mn ? umma::mnmajor_desc(sa, 4096) : kmajor_desc(sa);
        const uint64_t db = p.b_mn ? umma::mnmajor_desc(sa + A_BYTES, 4096) : kmajor_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < UK / 8; ++k) {
          const uint32_t accum = (i > 0 || k > 0) ? 1u : 0u;
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tmem), "l"(da + astep * k), "l"(db + bstep * k), "r"(idesc), "r"(accum) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tmem_full_bar)) : "memory");
    }
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) ----------------
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;
    if (nk > 0) {
      mbar_wait(&tmem_full_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;                      // warp-uniform
      uint32_t v[32];
      if (nk > 0) {
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int nb = n0 + c0;
      const bool row_ok = row < p.M;
      const bool full = (nb + 32 <= p.N);
      if (p.flags & GEMM_ATOMIC) {
        if (row_ok) {
          float* dst = p.C + (size_t)row * p.ldc + nb;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.N) atomicAdd(dst + j, __uint_as_float(v[j]));
          }
        }
        continue;
      }
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(v[j]);
        if (p.bias && (full || nb + j < p.N)) x += __ldg(p.bias + nb + j);
        f[j] = x;
      }
      if (p.C2 && row_ok) {
        float* dst = p.C2 + (size_t)row * p.ldc2 + nb;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) dst[j] = f[j];
        }
      }
      if (p.flags & GEMM_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (p.aux && row_ok) {
        const float* a = p.aux + (size_t)row * p.ldaux + nb;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(a + j));
            f[j] = m.x > 0.f ? f[j] : 0.f;
            f[j + 1] = m.y > 0.f ? f[j + 1] : 0.f;
            f[j + 2] = m.z > 0.f ? f[j + 2] : 0.f;
            f[j + 3] = m.w > 0.f ? f[j + 3] : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) f[j] = __ldg(a + j) > 0.f ? f[j] : 0.f;
        }
      }
      if (p.flags & GEMM_ROUND) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = round_tf32(f[j]);
      }
      if (p.C && row_ok) {
        float* dst = p.C + (size_t)row * p.ldc + nb;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) dst[j] = f[j];
        }
      }
      if (p.CT && row_ok) {   // transposed copy: for a fixed column the 32 lanes write 32 consecutive floats
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || nb + j < p.N) p.CT[(size_t)(nb + j) * p.ldct + row] = f[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(BN));
}

// mode 0 NN / 1 NT / 2 TN (see the file header).  Returns -3 for shapes the tcgen05 path does not take
// (MN-major dimensions must be multiples of 32, leading dimensions multiples of 4, 16-byte aligned bases);
// callers fall back to the mma.sync kernel (gemm_mma.cu) for those.
bool gemm_umma_supported(int mode, const GemmParams& g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  if ((g.lda & 3) || (g.ldb & 3) || ((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15)) return false;
  if (g.C && ((g.ldc & 3) || ((uintptr_t)g.C & 15))) return false;
  if (g.C2 && ((g.ldc2 & 3) || ((uintptr_t)g.C2 & 15))) return false;
  if (g.aux && ((g.ldaux & 3) || ((uintptr_t)g.aux & 15))) return false;
  if (mode == 2 && (g.M & 31)) return false;
  if (mode != 1 && (g.N & 31)) return false;
  return mode >= 0 && mode <= 2;
}

int gemm_umma(int mode, const GemmParams& g, float* CT, int ldct, int split_k, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return -1;
  if (!gemm_umma_supported(mode, g)) return -3;
  const int a_mn = (mode == 2), b_mn = (mode != 1);
  const int BN = g.N > 128 ? 256 : 128;
  CUtensorMap mA, mB;
  int rc = a_mn ? umma::make_map_blocks_mn(&mA, g.A, g.K, g.M, g.lda, UK, UM / 32) : make_map(&mA, g.A, g.M, g.K, g.lda, UM);
  if (rc) return rc;
  rc = b_mn ? umma::make_map_blocks_mn(&mB, g.B, g.K, g.N, g.ldb, UK, BN / 32) : make_map(&mB, g.B, g.N, g.K, g.ldb, BN);
  if (rc) return rc;
  UmmaParams p;
  p.C = g.C; p.ldc = g.ldc; p.CT = CT; p.ldct = ldct; p.C2 = g.C2; p.ldc2 = g.ldc2; p.bias = g.bias;
  p.aux = g.aux; p.ldaux = g.ldaux; p.M = g.M; p.N = g.N; p.K = g.K; p.flags = g.flags;
  p.a_mn = a_mn; p.b_mn = b_mn;
  int splits = (g.flags & GEMM_ATOMIC) ? (split_k > 0 ? split_k : 1) : 1;
  int kps = ((g.K + splits - 1) / splits + UK - 1) / UK * UK;
  splits = (g.K + kps - 1) / kps;
  p.k_per_split = kps;
  dim3 grid((g.N + BN - 1) / BN, (g.M + UM - 1) / UM, splits);
  const size_t smem = 1024 + (size_t)STAGES * (UM * UK * 4 + BN * UK * 4);
  if (BN == 256) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(gemm_umma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    gemm_umma_kernel<256><<<grid, THREADS, smem, st>>>(mA, mB, p);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(gemm_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    gemm_umma_kernel<128><<<grid, THREADS, smem, st>>>(mA, mB, p);
  }
  WN_CHECK_LAUNCH();
  return 0;
}

// C[M,N] (+)= A[M,K] . B[N,K]^T on tcgen05 (kept for the C ABI / tests).  GemmParams.B is the [N,K] operand here.
int gemm_nt_umma(const GemmParams& g, float* CT, int ldct, int split_k, cudaStream_t st) {
  return gemm_umma(1, g, CT, ldct, split_k, st);
}

// out[c][r] = in[r][c]  (32x32 tiles through shared memory, both sides coalesced)
__global__ void transpose_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int ldo, int rows,
                                 int cols, int round_out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ldi + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * ldo + r] = round_out ? round_tf32(tile[threadIdx.x][i]) : tile[threadIdx.x][i];
  }
}

int transpose(const float* in, int ldi, float* out, int ldo, int rows, int cols, int round_out, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return -1;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  if (grid.y > 65535) return -1;
  transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in, ldi, out, ldo, rows, cols, round_out);
  WN_CHECK_LAUNCH();
  return 0;
}

// out = tf32-rounded copy of in (tcgen05 kind::tf32 truncates raw fp32 operands; pre-rounding to
// nearest halves the error and removes its bias)
__global__ void round_copy_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = round_tf32(in[i]);
}
int round_copy(const float* in, float* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 2048) blocks = 2048;
  round_copy_kernel<<<(int)blocks, 256, 0, st>>>(in, out, n);
  WN_CHECK_LAUNCH();
  return 0;
}

}  // namespace wn
