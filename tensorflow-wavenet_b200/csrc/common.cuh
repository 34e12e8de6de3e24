// Shared device helpers for the sm_100a WaveNet kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#define WN_CHECK_LAUNCH()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

#define WN_BADARG(cond) \
  do {                  \
    if (cond) return -1; \
  } while (0)

namespace wn {

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// value of x rounded (to nearest, ties away) to the 10-bit tf32 mantissa, as a float
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float(f2tf32(x)); }

// D(16x8) += A(16x8,row) * B(8x8,col), tf32 inputs, fp32 accumulate.
// lane = 4*g + t.  A: a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4)
//                  B: b0=(k=t,n=g) b1=(k=t+4,n=g)
//                  C: c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 2.0f * sigmoid_f(2.0f * x) - 1.0f; }

// Same functions on the raw SFU instructions (ex2.approx.ftz / rcp.approx.ftz, ~2 ulp each, no range fix-up code):
// 2^(-x*log2e) overflows to +inf -> rcp gives 0, underflows to 0 -> rcp gives 1, both the right limits.
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.0f, sigmoid_fast(2.0f * x), -1.0f); }

// tanh(f) * sigmoid(g) = (1 - a) / ((1 + a)(1 + b)),  a = e^-2f, b = e^-g : two ex2 and ONE rcp (the SFU is the
// busiest unit of the block epilogue).  Arguments are clamped so that a, b stay finite (tanh and sigmoid are
// saturated to fp32 precision long before |f| = 20, |g| = 40).
__device__ __forceinline__ float gated_fast(float f, float g) {
  float a, b, r;
  f = fminf(fmaxf(f, -20.f), 20.f);
  g = fmaxf(g, -40.f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(-2.8853900817779268f * f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(-1.4426950408889634f * g));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((1.0f + a) * (1.0f + b)));
  return (1.0f - a) * r;
}

// tanh(f) and sigmoid(g) from the same three SFU operations (backward recompute)
__device__ __forceinline__ void gated_parts_fast(float f, float g, float& tf, float& sg) {
  float a, b, r;
  f = fminf(fmaxf(f, -20.f), 20.f);
  g = fmaxf(g, -40.f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(-2.8853900817779268f * f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(-1.4426950408889634f * g));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((1.0f + a) * (1.0f + b)));
  sg = r * (1.0f + a);
  tf = (1.0f - a) * (r * (1.0f + b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- programmatic dependent launch (PDL) ----
// The training step is one dependent chain of ~220 short kernels.  A kernel launched with launch_pdl() may
// become resident while its predecessor is still draining: everything before pdl_wait() (barrier init,
// TMEM allocation, the weight-image copy, tensor-map prefetch) overlaps the predecessor's tail; pdl_wait()
// returns when the predecessor grid has completed and its writes are visible.  Every kernel calls
// pdl_trigger() only AFTER its own pdl_wait(), so "my predecessor has triggered" implies "everything older
// than my predecessor has completed" -- data written two or more launches earlier may be read before the wait.
// A kernel triggers ONLY when the host tells it that its stream successor is such a waiting kernel (pdl_next):
// plain successors and event records behind a triggering kernel were observed to run / fire at the trigger,
// before the grid had finished (NaNs from half-written gradients).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static int no_pdl = -1;      // WN_NO_PDL=1: plain stream-ordered launches (debugging aid)
  if (no_pdl < 0) no_pdl = getenv("WN_NO_PDL") ? 1 : 0;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace wn
