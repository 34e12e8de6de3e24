// Helpers shared by the fp16 split-row residual-block kernels (block_fwd_h.cu: forward chain, block_bwd_h.cu: backward
// chain): tile constants, kind::f16 instruction / operand descriptors, tensor maps of split rows, TMA stores with L2
// hints, and the acquire / release tile flags that order the persistent kernels instead of kernel boundaries.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma_common.cuh"

namespace wn {
using namespace umma;
namespace {
constexpr int C = 32;
constexpr int TM = 128;
constexpr uint32_t TILE = TM * 128;                 // [128 rows][64 fp16]
constexpr uint32_t IMG_H = 8192 + 8192 + 4096;      // W0cat | W1cat | Wdcat

// byte offset of fp16 element (row, k) in a [rows][64 fp16] 128B-swizzled K-major tile
__device__ __forceinline__ uint32_t swzh(int row, int k) {
  return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7))) << 4) + ((k & 7) << 1));
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // fp16 x fp16 -> fp32, K-major
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the six (A k-step, B k-step) pairs of a split-precision product: hi.hi, lo.hi, hi.lo
__device__ __forceinline__ void mma_split(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool first) {
  mma_f16_ss(tmem_d, da + 0, db + 0, idesc, first ? 0u : 1u);
  mma_f16_ss(tmem_d, da + 2, db + 2, idesc, 1u);
  mma_f16_ss(tmem_d, da + 4, db + 0, idesc, 1u);
  mma_f16_ss(tmem_d, da + 6, db + 2, idesc, 1u);
  mma_f16_ss(tmem_d, da + 0, db + 4, idesc, 1u);
  mma_f16_ss(tmem_d, da + 2, db + 6, idesc, 1u);
}
__device__ __forceinline__ void split_h(float x, __half& h, __half& l) {
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}

static int make_map_split(CUtensorMap* m, const __half* ptr, int64_t B, int64_t T) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[3] = {64, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstr[2] = {128, (cuuint64_t)T * 128};
  cuuint32_t box[3] = {64, (cuuint32_t)TM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}
// fp16 copy of Zcat: [B][T][ldz] halfs, box = [TM rows][32 halfs] (64-byte rows, 64B swizzle)
static int make_map_z16(CUtensorMap* m, const __half* ptr, int64_t B, int64_t T, int64_t ldz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[3] = {(cuuint64_t)ldz, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)ldz * 2, (cuuint64_t)T * ldz * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)TM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// TMA store with an L2 eviction policy (createpolicy): the fp32 outputs of the forward stack are streamed out and not
// read again before the backward pass, the split rows are re-read by the next layer from the L2
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, const void* src, int c0, int c1, int c2, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }


__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const unsigned int* p) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    if (ld_acquire(p)) return;
    __nanosleep(32);
  }
  wait_timed_out(0xF1A6u, (unsigned int)(uintptr_t)p, gridDim.x, 0u);
}
// three flags at once (the loads overlap): all set?
__device__ __forceinline__ bool flags_set(const unsigned int* f0, const unsigned int* f1, const unsigned int* f2) {
  const unsigned int v0 = ld_acquire(f0), v1 = ld_acquire(f1), v2 = ld_acquire(f2);
  return (v0 & v1 & v2) != 0u;
}
__device__ __forceinline__ void wait_flags(const unsigned int* f0, const unsigned int* f1, const unsigned int* f2) {
  if (flags_set(f0, f1, f2)) return;
  wait_flag(f0);
  wait_flag(f1);
  wait_flag(f2);
}
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0u;
}

}  // namespace
}  // namespace wn
