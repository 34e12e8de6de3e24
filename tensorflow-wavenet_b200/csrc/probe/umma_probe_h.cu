// Stand-alone probe: tcgen05.mma kind::f16 with MN-major fp16 operands (the weight-gradient products read both operands
// with the NON-contracted dimension contiguous).  Operand tiles are loaded by TMA with the plain 128B swizzle as
// [K rows][64 halfs] boxes (128-byte rows); the probe tries the shared-memory descriptor conventions (LBO / SBO) and
// checks against exact CPU results (small dyadic inputs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe_h umma_probe_h.cu && ./umma_probe_h
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
  return (EncodeTiledFn)fn;
}
// fp16 row-major [rows][cols], box = [box_rows][64 halfs], 128B swizzle
static CUtensorMap make_map(EncodeTiledFn enc, const __half* ptr, int rows, int cols, int box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(2); }
  return m;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity, int max_spins) {
  for (int i = 0; i < max_spins; ++i) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
struct Case {
  int a_mn, b_mn;
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;
};
constexpr int M = 128, N = 128, K = 64;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Case c, float* __restrict__ C,
             int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;              // 16 KB
  unsigned char* sB = smem + 16384;      // 16 KB
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tma)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(32768) : "memory");
    auto load = [&](unsigned char* dst, const CUtensorMap* map, int c0, int c1) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(&bar_tma)), "r"(c0), "r"(c1) : "memory");
    };
    if (c.a_mn) { load(sA, &mapA, 0, 0); load(sA + 8192, &mapA, 64, 0); }      // two [64 k][64 m] boxes
    else load(sA, &mapA, 0, 0);                                                // one [128 m][64 k] box
    if (c.b_mn) { load(sB, &mapB, 0, 0); load(sB + 8192, &mapB, 64, 0); }
    else load(sB, &mapB, 0, 0);
    if (!mbar_wait(&bar_tma, 0, 4000000)) status[0] = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (status[0] == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(M >> 4) << 24);
      for (int k = 0; k < K / 16; ++k) {
        const uint64_t da = make_desc(smem_u32(sA) + k * c.a_kstep, c.a_lbo, c.a_sbo);
        const uint64_t db = make_desc(smem_u32(sB) + k * c.b_kstep, c.b_lbo, c.b_sbo);
        const uint32_t accum = k > 0 ? 1u : 0u;
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
  }
  __syncthreads();
  if (status[0] == 0 && !mbar_wait(&bar_mma, 0, 4000000)) status[0] = 2;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (status[0] == 0) {
    for (int n0 = 0; n0 < N; n0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + n0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = warp * 32 + lane;
      for (int j = 0; j < 16; ++j) C[row * N + n0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

static float dy(int i) { return (float)((i * 37 + 11) % 17 - 8) / 8.0f; }

int main() {
  EncodeTiledFn enc = get_encode();
  // logical A[m][k], B[n][k]
  std::vector<float> A(M * K), B(N * K), ref(M * N);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m * K + k] = dy(m * 131 + k * 7);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n * K + k] = dy(n * 53 + k * 29 + 5);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    float s = 0;
    for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
    ref[m * N + n] = s;
  }
  std::vector<__half> Akm(M * K), Amn(K * M), Bkm(N * K), Bmn(K * N);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { Akm[m * K + k] = __float2half(A[m * K + k]); Amn[k * M + m] = Akm[m * K + k]; }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { Bkm[n * K + k] = __float2half(B[n * K + k]); Bmn[k * N + n] = Bkm[n * K + k]; }
  __half *dAkm, *dAmn, *dBkm, *dBmn;
  float* dC;
  int* dS;
  CK(cudaMalloc(&dAkm, M * K * 2)); CK(cudaMalloc(&dAmn, M * K * 2)); CK(cudaMalloc(&dBkm, N * K * 2)); CK(cudaMalloc(&dBmn, N * K * 2));
  CK(cudaMalloc(&dC, M * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dAkm, Akm.data(), M * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAmn, Amn.data(), M * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBkm, Bkm.data(), N * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBmn, Bmn.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
  const uint32_t conv[4][3] = {{8192, 1024, 2048}, {1024, 8192, 2048}, {8192, 2048, 2048}, {2048, 8192, 2048}};   // {LBO, SBO, k step}
  int n_fail = 0, n_pass = 0;
  for (int which = 0; which < 4; ++which)        // 0 both K-major, 1 A MN, 2 B MN, 3 both MN
    for (int ci = 0; ci < (which ? 4 : 1); ++ci) {
      Case c;
      c.a_mn = (which == 1 || which == 3); c.b_mn = (which == 2 || which == 3);
      c.a_lbo = c.a_mn ? conv[ci][0] : 16; c.a_sbo = c.a_mn ? conv[ci][1] : 1024; c.a_kstep = c.a_mn ? conv[ci][2] : 32;
      c.b_lbo = c.b_mn ? conv[ci][0] : 16; c.b_sbo = c.b_mn ? conv[ci][1] : 1024; c.b_kstep = c.b_mn ? conv[ci][2] : 32;
      CUtensorMap mA = c.a_mn ? make_map(enc, dAmn, K, M, 64) : make_map(enc, dAkm, M, K, 128);
      CUtensorMap mB = c.b_mn ? make_map(enc, dBmn, K, N, 64) : make_map(enc, dBkm, N, K, 128);
      CK(cudaMemset(dC, 0, M * N * 4)); CK(cudaMemset(dS, 0, 4));
      probe_kernel<<<1, 128, 33 * 1024>>>(mA, mB, c, dC, dS);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> out(M * N);
      int st = -1;
      if (e == cudaSuccess) { CK(cudaMemcpy(out.data(), dC, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost)); }
      int nbad = 0;
      for (int i = 0; i < M * N; ++i) if (out[i] != ref[i]) ++nbad;
      const bool pass = (e == cudaSuccess && st == 0 && nbad == 0);
      printf("[%s] A %s, B %s, MN-major {LBO %u, SBO %u, kstep %u} | cuda=%s status=%d mismatches=%d/%d\n", pass ? "PASS" : "FAIL",
             c.a_mn ? "MN" : "K", c.b_mn ? "MN" : "K", conv[ci][0], conv[ci][1], conv[ci][2], cudaGetErrorString(e), st, nbad, M * N);
      if (pass) ++n_pass; else ++n_fail;
      if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
    }
  printf("probe done: %d passed, %d failed\n", n_pass, n_fail);
  return 0;
}
