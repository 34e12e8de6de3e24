// Stand-alone probe of the sm_100a tensor-core path used by the production kernels:
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma kind::tf32 ->
// TMEM -> tcgen05.ld.  It checks, against exact CPU results (inputs are small dyadic numbers,
// so TF32 products and fp32 sums are exact), every operand form the GEMM / block kernels rely
// on: K-major A and B, MN-major A and B (shared-memory descriptor LBO/SBO conventions), and an
// A tile written by threads with the software swizzle + fence.proxy.async.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
// 2-D fp32 row-major [rows][cols] tensor, box = [box_rows][32 floats], 128B swizzle
static CUtensorMap make_map(EncodeTiledFn enc, const float* ptr, int rows, int cols, int box_rows,
                            CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(2); }
  return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity, int max_spins) {
  for (int i = 0; i < max_spins; ++i) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

struct Case {
  int a_mn, b_mn;            // operand majors (0 = K-major, 1 = MN-major)
  int a_boxes, a_box_rows, a_box_dc, a_box_dr;   // TMA boxes of A: count, rows per box, coordinate steps
  int b_boxes, b_box_rows, b_box_dc, b_box_dr;
  uint32_t a_lbo, a_sbo, a_kstep;                // descriptor fields (bytes) and start-address step per k-step
  uint32_t b_lbo, b_sbo, b_kstep;
  int N, ksteps, manual_a;
  int a_lt, b_lt;            // descriptor layout types (0 -> default 2)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Case c,
             const float* __restrict__ Araw, float* __restrict__ C, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);              // 16 KB
  float* sB = reinterpret_cast<float*>(smem + 16384);      // up to 32 KB
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (c.manual_a) {
    // threads write the K-major A tile [128][32] with the 128B software swizzle:
    // byte offset = r*128 + (((c/4) ^ (r%8)) * 16) + (c%4)*4
    for (int i = threadIdx.x; i < 128 * 32; i += 128) {
      const int r = i >> 5, col = i & 31;
      const int off = r * 128 + ((((col >> 2) ^ (r & 7)) << 4)) + (col & 3) * 4;
      *reinterpret_cast<float*>(smem + off) = Araw[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  int ok = 1;
  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
    if (!c.manual_a) bytes += c.a_boxes * c.a_box_rows * 128;
    bytes += c.b_boxes * c.b_box_rows * 128;
    mbar_expect_tx(&bar_tma, bytes);
    if (!c.manual_a)
      for (int i = 0; i < c.a_boxes; ++i)
        tma_load_2d(reinterpret_cast<unsigned char*>(sA) + i * c.a_box_rows * 128, &mapA, &bar_tma, i * c.a_box_dc, i * c.a_box_dr);
    for (int i = 0; i < c.b_boxes; ++i)
      tma_load_2d(reinterpret_cast<unsigned char*>(sB) + i * c.b_box_rows * 128, &mapB, &bar_tma, i * c.b_box_dc, i * c.b_box_dr);
    if (!mbar_wait(&bar_tma, 0, 4000000)) { ok = 0; status[0] = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
      uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                       ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int k = 0; k < c.ksteps; ++k) {
        const uint64_t da = make_desc(smem_u32(sA) + k * c.a_kstep, c.a_lbo, c.a_sbo, c.a_lt ? c.a_lt : 2);
        const uint64_t db = make_desc(smem_u32(sB) + k * c.b_kstep, c.b_lbo, c.b_sbo, c.b_lt ? c.b_lt : 2);
        const uint32_t accum = k > 0 ? 1u : 0u;
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
  }
  __syncthreads();
  // everybody waits for the MMA completion (bounded)
  if (status[0] == 0) {
    if (!mbar_wait(&bar_mma, 0, 4000000)) { status[0] = 2; }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (status[0] == 0) {
    for (int n0 = 0; n0 < c.N; n0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + n0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = warp * 32 + lane;
      for (int j = 0; j < 32; ++j) C[row * c.N + n0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static float dy(int i) { return (float)((i * 37 + 11) % 17 - 8) / 8.0f; }   // dyadic, exact in tf32

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d\n", prop.name, prop.major, prop.minor);
  EncodeTiledFn enc = get_encode();
  const int M = 128, K = 32;
  int n_fail = 0;
  for (int variant = 0; variant < 20; ++variant) {
    Case c;
    memset(&c, 0, sizeof(c));
    const char* name = "";
    int N = 64;
    // defaults: K-major A via one box [128 x 32], K-major B via one box [N x 32]
    c.a_boxes = 1; c.a_box_rows = 128; c.a_lbo = 16; c.a_sbo = 1024; c.a_kstep = 32;
    c.b_boxes = 1; c.b_box_rows = N; c.b_lbo = 16; c.b_sbo = 1024; c.b_kstep = 32;
    c.ksteps = 4;
    switch (variant) {
      case 0: name = "A K-major, B K-major (N=64)"; break;
      case 1: name = "A K-major, B K-major (N=256)"; N = 256; c.b_box_rows = 256; break;
      case 2: name = "A K-major, B MN-major LBO=4096 SBO=1024"; c.b_mn = 1; c.b_boxes = 2; c.b_box_rows = 32; c.b_box_dc = 32;
              c.b_lbo = 4096; c.b_sbo = 1024; c.b_kstep = 1024; break;
      case 3: name = "A K-major, B MN-major LBO=1024 SBO=4096 (swapped)"; c.b_mn = 1; c.b_boxes = 2; c.b_box_rows = 32; c.b_box_dc = 32;
              c.b_lbo = 1024; c.b_sbo = 4096; c.b_kstep = 1024; break;
      case 4: name = "A MN-major LBO=4096 SBO=1024, B K-major"; c.a_mn = 1; c.a_boxes = 4; c.a_box_rows = 32; c.a_box_dc = 32;
              c.a_lbo = 4096; c.a_sbo = 1024; c.a_kstep = 1024; break;
      case 5: name = "A MN-major LBO=1024 SBO=4096 (swapped), B K-major"; c.a_mn = 1; c.a_boxes = 4; c.a_box_rows = 32; c.a_box_dc = 32;
              c.a_lbo = 1024; c.a_sbo = 4096; c.a_kstep = 1024; break;
      case 6: name = "A K-major written by threads (software swizzle + fence.proxy.async), B K-major"; c.manual_a = 1; break;
      case 7: name = "A MN-major, B MN-major (N=128)"; N = 128; c.a_mn = 1; c.a_boxes = 4; c.a_box_rows = 32; c.a_box_dc = 32;
              c.a_lbo = 4096; c.a_sbo = 1024; c.a_kstep = 1024;
              c.b_mn = 1; c.b_boxes = 4; c.b_box_rows = 32; c.b_box_dc = 32; c.b_lbo = 4096; c.b_sbo = 1024; c.b_kstep = 1024; break;
    }
    // ---- tf32 MN-major operands: 128B swizzle with 32-byte atoms (TMA SWIZZLE_128B_ATOM_32B, descriptor layout 1)
    char namebuf[160];
    if (variant >= 8) {
      static const uint32_t cand[6][2] = {{4096, 512}, {512, 4096}, {4096, 1024}, {1024, 4096}, {4096, 256}, {256, 4096}};
      const int which = (variant - 8) / 6;        // 0: B MN-major, 1: A MN-major
      const int ci = (variant - 8) % 6;
      if (which > 1) break;
      if (which == 0) { c.b_mn = 1; c.b_lt = 1; c.b_boxes = 2; c.b_box_rows = 32; c.b_box_dc = 32; c.b_lbo = cand[ci][0]; c.b_sbo = cand[ci][1]; c.b_kstep = 1024; }
      else { c.a_mn = 1; c.a_lt = 1; c.a_boxes = 4; c.a_box_rows = 32; c.a_box_dc = 32; c.a_lbo = cand[ci][0]; c.a_sbo = cand[ci][1]; c.a_kstep = 1024; }
      snprintf(namebuf, sizeof(namebuf), "%s MN-major SW128_ATOM_32B layout=1 LBO=%u SBO=%u", which ? "A" : "B", cand[ci][0], cand[ci][1]);
      name = namebuf;
    }
    c.N = N;
    // logical matrices
    std::vector<float> A(M * K), B(K * N), Cref(M * N), Cout(M * N, -777.f);
    for (int i = 0; i < M * K; ++i) A[i] = dy(i);
    for (int i = 0; i < K * N; ++i) B[i] = dy(i * 3 + 5);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float s = 0;
        for (int k = 0; k < K; ++k) s += A[m * K + k] * B[k * N + n];
        Cref[m * N + n] = s;
      }
    // storage: K-major A = [M][K]; MN-major A = [K][M]; K-major B = [N][K]; MN-major B = [K][N]
    std::vector<float> As(M * K), Bs(K * N);
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) As[c.a_mn ? k * M + m : m * K + k] = A[m * K + k];
    for (int k = 0; k < K; ++k)
      for (int n = 0; n < N; ++n) Bs[c.b_mn ? k * N + n : n * K + k] = B[k * N + n];
    float *dA, *dB, *dC;
    int* dS;
    CK(cudaMalloc(&dA, As.size() * 4)); CK(cudaMalloc(&dB, Bs.size() * 4)); CK(cudaMalloc(&dC, Cout.size() * 4));
    CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, As.data(), As.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bs.data(), Bs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, Cout.data(), Cout.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dS, 0, 4));
    CUtensorMap mA = c.a_mn ? make_map(enc, dA, K, M, c.a_box_rows, c.a_lt == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B) : make_map(enc, dA, M, K, c.a_box_rows);
    CUtensorMap mB = c.b_mn ? make_map(enc, dB, K, N, c.b_box_rows, c.b_lt == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B) : make_map(enc, dB, N, K, c.b_box_rows);
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024));
    probe_kernel<<<1, 128, 16384 + 32768 + 1024>>>(mA, mB, c, dA, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = -1;
    if (e == cudaSuccess) {
      CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(Cout.data(), dC, Cout.size() * 4, cudaMemcpyDeviceToHost));
    }
    double maxerr = 0;
    int nbad = 0;
    for (int i = 0; i < M * N; ++i) {
      double d = fabs((double)Cout[i] - (double)Cref[i]);
      if (d > maxerr) maxerr = d;
      if (d != 0) ++nbad;
    }
    const bool pass = (e == cudaSuccess && st == 0 && nbad == 0);
    printf("[%s] variant %d: %s | cuda=%s status=%d mismatches=%d/%d maxerr=%g  C[0..3]=%g %g %g %g ref %g %g %g %g\n",
           pass ? "PASS" : "FAIL", variant, name, cudaGetErrorString(e), st, nbad, M * N, maxerr, Cout[0], Cout[1], Cout[2],
           Cout[3], Cref[0], Cref[1], Cref[2], Cref[3]);
    if (!pass) ++n_fail;
    if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 3; }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dS);
  }
  printf("probe done, %d variant(s) failed\n", n_fail);
  return 0;
}
