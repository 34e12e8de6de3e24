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
// Structure (persistent CTAs, one per SM, looping over 128 x BN output tiles x K splits, BN in {128, 256}):
//   warp 0 : TMA producer   -- cp.async.bulk.tensor 2D boxes [128|BN rows x 32 floats] into a
//            4-stage ring of 128B-swizzled tiles, completion on "full" mbarriers
//   warp 1 : MMA issuer     -- one lane issues 4 x tcgen05.mma.kind::tf32 (K=8 each) per stage into one of
//            TWO 128 x BN fp32 accumulators in TMEM; tcgen05.commit releases the stage ("empty" mbarrier)
//   warps 2-5 : epilogue    -- tcgen05.ld the accumulator (each warp its TMEM lane quadrant), apply
//            bias / relu / relu-gradient mask / tf32 rounding, store row-major and (optionally) a
//            transposed copy (coalesced: lanes hold consecutive rows), or red.add for split-K.
// Descriptor conventions were validated on B200 by probe/umma_probe.cu.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "umma_common.cuh"

namespace wn {

namespace {

constexpr int UM = 128;        // UMMA M (rows of A per tile)
constexpr int UK = 32;         // floats per stage along K (= one 128-byte swizzle row)
constexpr int MAX_STAGES = 6;        // ring depth: 3-4 (48 KB stages), up to 6 for CTA pairs (32 KB stages)
constexpr uint32_t STG_BYTES = 4096;   // one [32 rows][32 floats] epilogue staging block (128B-swizzled)
constexpr int THREADS = 320;      // TMA warp, MMA warp, up to 8 epilogue warps (p.ewarps = 4 or 8)

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

// fp16 row-major [rows][cols] (row pitch ld halfs); box = [box_rows][box_cols halfs]
int make_map16(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
               CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
#pragma unroll 1
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
  int k_per_split;                // K range per split (multiple of UK)
  int splits;
  int stages;                     // depth of the TMA -> MMA ring
  int a_mn, b_mn;                 // operand majors: 0 = K contiguous, 1 = M / N contiguous
  int f16;                        // operands are fp16 (K-major, 64 elements per stage, kind::f16); C stays fp32
  int has_c16;                    // additional fp16 copy of the accumulators through mapC16 (the next GEMM's A operand)
  uint32_t* mask_out;             // relu mask of this product as bits, [M][ldmw] words (bit j of word w: column 32 w + j)
  const uint32_t* mask_in;        // zero the outputs whose bit is clear (the relu mask of the forward product)
  int ldmw;
  int m2;                         // 256-row tiles: two M=128 MMAs per k-step share the B operand (split-K weight-gradient form only:
                                  // both TMEM accumulators belong to one tile, no epilogue overlap)
  int ewarps;                     // epilogue warps: 4 (a warp takes all BN columns of its 32 rows) or 8 (two warps per row quadrant, BN/2 columns each)
  int nbuf;                       // depth of the per-warp output staging rings (2..4 TMA stores in flight per warp)
  float c_scale;                  // C = c_scale * (accumulator, bias, relu, mask); the fp16 copy stays unscaled (0: no scaling)
  float* colsum;                  // [N] += colsum_scale * column sums of the fp16 copy (needs has_c16 and no bias: the sums
  float colsum_scale;             // collect in the bias row of shared memory); the bias gradient of the layer below
  // dilated (two-tap) A operand of the wide residual blocks (block_wide16.cu): K chunks below a_split come from rows
  // m + a_shift of the SAME [rows][a_split] matrix, chunks from a_split on from rows m (columns k - a_split); rows outside
  // the matrix are zero-filled by the TMA unit = the causal padding.  0: plain A[M][K]
  int a_split, a_shift;
  int tn_R, tn_shift;             // two-tap weight-gradient form: output rows [0, tn_R) pair A row k with B row k + tn_shift, rows
                                  // [tn_R, 2 tn_R) (A columns m - tn_R) pair row k with row k:  [x[t-d] | x[t]]^T . dpre in one launch
  int wstage;                     // fp16-only outputs, 8 epilogue warps: a warp stages ALL its columns of a tile ([32 rows][64 halfs] blocks,
                                  // 128B swizzle) and issues its TMA stores once per tile; the accumulator is released before the stores
  int dpre;                       // N = BN = 128 (block_wide16.cu backward): the product + aux is dz; the epilogue loads the saved
                                  // pre-activations [f 128 | g 128] through mapC and stores dpre = [df | dg] ([M][256]) through mapC16
  int gate;                       // N = BN = 256 = [f | g]: z = tanh(f + b) sigmoid(g + b) stored through mapC (fp16), f / g through mapC16
  int aux_add;                    // aux is an fp16 matrix ADDED (times aux_scale) to the accumulator (residual / skip-path term)
  float aux_scale;
};

// PAIR = 1: CTA pairs (clusters of 2, tcgen05 cta_group::2): one M = 256 MMA per k-step spans both CTAs of a pair; a CTA
// loads its own 128 rows of A and HALF of the B tile, so an SM receives 32 KB instead of 48 KB per k-step (BN = 256) -- the
// long-K products are bound by exactly that L2 -> SM traffic.  The even CTA (leader) issues the MMAs and commits to the
// barriers of both; TMA loads of both CTAs complete on the leader's "full" barrier; each CTA runs the epilogue of its own
// 128 rows out of its own TMEM.  fp16 NT form only.
template <int BN, int PAIR>
__global__ void __launch_bounds__(THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapAux, const __grid_constant__ CUtensorMap mapC16,
                 UmmaParams p) {
  const uint32_t A_BYTES = (p.m2 ? 2 : 1) * UM * UK * 4;       // 16 KB (32 KB for 256-row tiles)
  constexpr uint32_t B_BYTES = BN * UK * 4 / (PAIR ? 2 : 1);   // 16 / 32 KB (a pair: each CTA holds half of the B tile)
  const uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  const int TM_ROWS = PAIR ? 2 * UM : (p.m2 ? 2 * UM : UM);    // rows of an output tile (a pair: 128 per CTA)
  uint32_t cta_rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // persistent loop over items: per CTA / per pair
  const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tmem_full_bar[2], tmem_empty_bar[2], aux_bar[8][2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[256];      // bias of the tile's columns / collected column sums (dpre mode: 256 of them)
  // epilogue staging (per epilogue warp, double buffered): output chunks leave through TMA stores, the
  // relu-gradient mask chunks arrive through TMA loads -- no scattered 16-byte global accesses
  const int STAGES = p.stages;
  unsigned char* out_stage = smem + STAGES * STAGE_BYTES;
  // staging regions exist only for what the launch uses (same arithmetic as staging_bytes() on the host)
  unsigned char* aux_stage = out_stage + (p.C ? p.ewarps * p.nbuf * STG_BYTES : 0);         // [warps][nbuf][32 rows][32 floats]
  unsigned char* c16_stage = aux_stage + (p.aux ? p.ewarps * 2 * STG_BYTES : 0);            // [warps][nbuf][32 rows][32 halfs]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Persistent CTA: work items (output tile x K split), N tile fastest so that CTAs running side by side share
  // the A tile in L2.  Two accumulators in TMEM: the epilogue of item i overlaps the main loop of item i+1.
  const int n_nt = (p.N + BN - 1) / BN, n_mt = (p.M + TM_ROWS - 1) / TM_ROWS;
  const int n_items = n_nt * n_mt * p.splits;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], PAIR ? 2 * p.ewarps : p.ewarps); }
    for (int s = 0; s < 16; ++s) mbar_init(&aux_bar[s >> 1][s & 1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * BN));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * BN));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) {      // both CTAs' barriers exist before either signals the other
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const int ukk = p.f16 ? 2 * UK : UK;      // K elements per stage: one 128-byte row of either type
  auto decode = [&](int item, int& m0, int& n0, int& k_begin, int& nk) {
    const int nt = item % n_nt;
    const int rest = item / n_nt;
    const int mt = rest % n_mt, sp = rest / n_mt;
    m0 = mt * TM_ROWS + (PAIR ? (int)cta_rank * UM : 0);      // (a pair: this CTA's 128 rows)
    n0 = nt * BN;
    k_begin = sp * p.k_per_split;
    int k_end = k_begin + p.k_per_split;
    if (k_end > p.K) k_end = p.K;
    nk = (k_end - k_begin + ukk - 1) / ukk;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      uint32_t lead_full[MAX_STAGES];      // a pair: the leader's "full" barriers (cluster addresses)
      if (PAIR)
        for (int s = 0; s < MAX_STAGES; ++s)
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lead_full[s]) : "r"(smem_u32(&full_bar[s])), "r"(0));
      for (int item = item0; item < n_items; item += item_step) {
        int m0, n0, k_begin, nk;
        decode(item, m0, n0, k_begin, nk);
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* sa = smem + s * STAGE_BYTES;
          const int kc = k_begin + i * ukk;
          if (PAIR) {      // both CTAs' boxes complete on the leader's barrier, which expects the bytes of both
            if (cta_rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
            if (p.a_mn) {      // weight-gradient form: MN-major blocks of this CTA's 128 A columns and its half of the B columns
              const int tap1 = (p.tn_R && m0 >= p.tn_R) ? 1 : 0;
              const int kr = kc - ((p.tn_R && !tap1) ? p.tn_shift : 0);      // two-tap: the past tap reads A rows k - d (B rows are shared)
              asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                           ::"r"(smem_u32(sa)), "l"(&mapA), "r"(lead_full[s]), "r"(0), "r"(kr), "r"((m0 - tap1 * p.tn_R) >> 6) : "memory");
              asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                           ::"r"(smem_u32(sa + A_BYTES)), "l"(&mapB), "r"(lead_full[s]), "r"(0), "r"(kc), "r"((n0 + (int)cta_rank * (BN / 2)) >> 6) : "memory");
              continue;
            }
            const int ka = (p.a_split && kc < p.a_split) ? kc : kc - p.a_split;
            const int ma = (p.a_split && kc < p.a_split) ? m0 + p.a_shift : m0;
            asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(sa)), "l"(&mapA), "r"(lead_full[s]), "r"(ka), "r"(ma) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(sa + A_BYTES)), "l"(&mapB), "r"(lead_full[s]), "r"(kc), "r"(n0 + (int)cta_rank * (BN / 2)) : "memory");
            continue;
          }
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          const int bsh = p.f16 ? 6 : 5;      // MN-major blocks: 32 floats / 64 halfs wide
          const int tap1 = (p.tn_R && m0 >= p.tn_R) ? 1 : 0;
          const int kr = kc - ((p.tn_R && !tap1) ? p.tn_shift : 0);      // two-tap: the past tap pairs A row k - d with B row k
          const int kb = kc;
          if (p.a_mn) umma::tma_load_3d(sa, &mapA, &full_bar[s], 0, kr, (m0 - tap1 * p.tn_R) >> bsh);     // [blocks][k rows][128 B of m]
          else if (p.a_split && kc < p.a_split) tma_load_2d(sa, &mapA, &full_bar[s], kc, m0 + p.a_shift);
          else tma_load_2d(sa, &mapA, &full_bar[s], kc - p.a_split, m0);
          if (p.b_mn) umma::tma_load_3d(sa + A_BYTES, &mapB, &full_bar[s], 0, kb, n0 >> bsh);
          else tma_load_2d(sa + A_BYTES, &mapB, &full_bar[s], kc, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = p.f16 ? ((1u << 4) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                                      ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * UM : UM) >> 4) << 24))
                                   : umma::idesc_tf32(UM, BN, p.a_mn, p.b_mn);
      // descriptor start-address step per MMA (>> 4): K-major 32 B; MN-major tf32 8 rows x 128 B, fp16 16 rows x 128 B
      const uint32_t mnstep = p.f16 ? 128 : 64;
      const uint32_t astep = p.a_mn ? mnstep : 2, bstep = p.b_mn ? mnstep : 2;
      // MN-major fp16 (probe/umma_probe_h.cu): plain 128B swizzle, blocks of 64 elements x 64 k rows (LBO 8192), SBO 1024
      auto mn_desc = [&](uint32_t saddr) -> uint64_t {
        if (!p.f16) return umma::mnmajor_desc(saddr, 4096);
        uint64_t d = 0;
        d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
        d |= (uint64_t)(8192 >> 4) << 16;
        d |= (uint64_t)(1024 >> 4) << 32;
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)2 << 61;
        return d;
      };
      uint32_t it = 0, local = 0;
      for (int item = item0; item < n_items; item += item_step, ++local) {
        int m0, n0, k_begin, nk;
        decode(item, m0, n0, k_begin, nk);
        const uint32_t acc = p.m2 ? 0u : (local & 1), aph = p.m2 ? (local & 1) : ((local >> 1) & 1);
        mbar_wait(&tmem_empty_bar[acc], aph ^ 1);      // the epilogue (a pair: of both CTAs) has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem + acc * BN;
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = p.a_mn ? mn_desc(sa) : kmajor_desc(sa);
          const uint64_t db = p.b_mn ? mn_desc(sa + A_BYTES) : kmajor_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < UK / 8; ++k) {
            const uint32_t accum = (i > 0 || k > 0) ? 1u : 0u;
            if (PAIR) {
              asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                           ::"r"(d_tmem), "l"(da + astep * k), "l"(db + bstep * k), "r"(idesc), "r"(accum) : "memory");
            } else if (p.f16) {
              asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                           ::"r"(d_tmem), "l"(da + astep * k), "l"(db + bstep * k), "r"(idesc), "r"(accum) : "memory");
              if (p.m2)      // rows 128..255 of the tile: the second half of the A stage (+16 KB), second accumulator
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                             ::"r"(d_tmem + BN), "l"(da + 1024 + astep * k), "l"(db + bstep * k), "r"(idesc), "r"(accum) : "memory");
            } else
              asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                           ::"r"(d_tmem), "l"(da + astep * k), "l"(db + bstep * k), "r"(idesc), "r"(accum) : "memory");
          }
          if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(smem_u32(&empty_bar[s])), "h"((uint16_t)3) : "memory");      // the stage is free in BOTH CTAs
          else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
        }
        if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                               ::"r"(smem_u32(&tmem_full_bar[acc])), "h"((uint16_t)3) : "memory");
        else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tmem_full_bar[acc])) : "memory");
      }
    }
  } else if (warp - 2 < p.ewarps) {
    // ---------------- epilogue: warps 2.. own TMEM lane quadrants (warp % 4) ----------------
    // One warp per scheduler cannot hide its own instruction latencies: measured, the epilogue (not the MMA pipe, L2
    // or HBM) set the pace of every GEMM with a short K loop at ~17.6 k cycles per 128 x 256 tile.  With 8 epilogue
    // warps two warps share a TMEM lane quadrant (lanes 32 (warp % 4) ...) and split the tile's columns.
    const int quad = warp & 3;
    const int ew = warp - 2;                                 // 0 .. ewarps-1
    const int etid = threadIdx.x - 64;                       // 0 .. 32 ewarps - 1 among the epilogue warps
    const int ethreads = 32 * p.ewarps;
    const int col_lo = (p.ewarps == 8 && ew >= 4) ? BN / 2 : 0;
    const int col_hi = (p.ewarps == 8 && ew < 4) ? BN / 2 : BN;
    unsigned char* my_out = out_stage + ew * p.nbuf * STG_BYTES;
    unsigned char* my_aux = aux_stage + ew * 2 * STG_BYTES;
    unsigned char* my_c16 = c16_stage + ew * p.nbuf * (STG_BYTES / 2);
    const bool staged = !(p.flags & GEMM_ATOMIC) && (p.C != nullptr || p.has_c16);
    uint32_t local = 0, out_cnt = 0, aux_cnt = 0;
    int cs_n0 = -1;      // column range whose sums bias_s currently holds
    auto colsum_flush = [&](int next_n0) {      // all epilogue warps: add the collected sums to global memory, start over
      asm volatile("bar.sync 1, %0;" ::"r"(ethreads) : "memory");
      for (int i = etid; i < (p.dpre ? 256 : BN); i += ethreads) {
        if (cs_n0 >= 0 && cs_n0 + i < (p.dpre ? 256 : p.N)) atomicAdd(p.colsum + cs_n0 + i, bias_s[i] * p.colsum_scale);
        bias_s[i] = 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(ethreads) : "memory");
      cs_n0 = next_n0;
    };
    // (a pair: every epilogue warp of both CTAs arrives on the LEADER's "accumulator drained" barrier)
    auto release_acc = [&](uint32_t acc_) {
      if (PAIR) {
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&tmem_empty_bar[acc_])), "r"(0));
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc_])) : "memory");
      }
    };
    for (int item = item0; item < n_items; item += item_step, ++local) {
      int m0, n0, k_begin, nk;
      decode(item, m0, n0, k_begin, nk);
      if (p.colsum && n0 != cs_n0) colsum_flush(n0);      // (warp-uniform, CTA-uniform)
      const uint32_t acc = p.m2 ? 0u : (local & 1), aph = p.m2 ? (local & 1) : ((local >> 1) & 1);
      const int row0 = m0 + quad * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      if (p.bias) {   // bias of this tile's columns -> shared memory (all four epilogue warps)
        asm volatile("bar.sync 1, %0;" ::"r"(ethreads) : "memory");     // previous tile's readers are done
        for (int i = etid; i < BN; i += ethreads) bias_s[i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"r"(ethreads) : "memory");
      }
      auto issue_aux = [&](int c0) {   // lane 0: mask chunk [32 rows][32 cols] -> staging buffer (aux_cnt parity)
        const uint32_t buf = (aux_cnt + (c0 > col_lo ? 1u : 0u)) & 1u;
        mbar_expect_tx(&aux_bar[ew][buf], p.aux_add ? STG_BYTES / 2 : STG_BYTES);
        tma_load_2d(my_aux + buf * STG_BYTES, &mapAux, &aux_bar[ew][buf], n0 + c0, row0);
      };
      if (p.aux && !p.wstage && !p.dpre && lane == 0 && n0 + col_lo < p.N) issue_aux(col_lo);
      uint32_t mw[8];      // this row's mask words of the tile, fetched before the accumulator is waited for
#pragma unroll
      for (int q = 0; q < 8; ++q) mw[q] = 0xFFFFFFFFu;
      if (p.mask_in && row_ok) {
#pragma unroll
        for (int q = 0; q < BN / 32; ++q)
          if (n0 + 32 * q < p.N) mw[q] = __ldg(p.mask_in + (size_t)row * p.ldmw + (n0 >> 5) + q);
      }
      auto tld = [&](uint32_t (&v)[32], int c0) {
        const uint32_t taddr = tmem + acc * BN + ((uint32_t)(quad * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
      };
      if (BN == 128 && p.dpre) {
        // ---- dz -> dpre in the epilogue: a warp takes 64 dz columns of its 32 rows, the matching saved f and g blocks and
        // the skip-path gradient block arrive by TMA while the MMAs run; df / dg leave as two blocks, their column sums
        // (bias / conditioning gradients) are collected on the way ----
        unsigned char* ds = out_stage + ew * 5 * 4096;        // blocks: dz_skip, f, g (in), df, dg (out)
        const int cf = (ew >> 2) * 64;
        if (lane == 0) {
          mbar_expect_tx(&aux_bar[ew][0], 3 * 4096);
          tma_load_2d(ds, &mapAux, &aux_bar[ew][0], cf, row0);
          tma_load_2d(ds + 4096, &mapC, &aux_bar[ew][0], cf, row0);
          tma_load_2d(ds + 8192, &mapC, &aux_bar[ew][0], 128 + cf, row0);
        }
        mbar_wait(&tmem_full_bar[acc], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        mbar_wait(&aux_bar[ew][0], aux_cnt & 1u);
        ++aux_cnt;
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t va[32];
          tld(va, cf + 32 * sub);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const unsigned char* is = ds + lane * 128;
          uint4 qf[4], qg[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t off = (uint32_t)((sub * 4 + j) ^ (lane & 7)) << 4;
            const uint4 qs = *reinterpret_cast<const uint4*>(is + off);
            const uint4 pf = *reinterpret_cast<const uint4*>(is + 4096 + off);
            const uint4 pg = *reinterpret_cast<const uint4*>(is + 8192 + off);
            const __half2* hs = reinterpret_cast<const __half2*>(&qs);
            const __half2* hf = reinterpret_cast<const __half2*>(&pf);
            const __half2* hg = reinterpret_cast<const __half2*>(&pg);
            uint32_t* of = reinterpret_cast<uint32_t*>(&qf[j]);
            uint32_t* og = reinterpret_cast<uint32_t*>(&qg[j]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 sk = __half22float2(hs[u]), ff = __half22float2(hf[u]), gg = __half22float2(hg[u]);
              const float dz0 = fmaf(p.aux_scale, sk.x, __uint_as_float(va[8 * j + 2 * u]));
              const float dz1 = fmaf(p.aux_scale, sk.y, __uint_as_float(va[8 * j + 2 * u + 1]));
              float t0, s0, t1, s1;
              gated_parts_fast(ff.x, gg.x, t0, s0);
              gated_parts_fast(ff.y, gg.y, t1, s1);
              const float df0 = dz0 * s0 * (1.f - t0 * t0), df1 = dz1 * s1 * (1.f - t1 * t1);
              const float dg0 = dz0 * t0 * s0 * (1.f - s0), dg1 = dz1 * t1 * s1 * (1.f - s1);
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(of[u]) : "f"(df1), "f"(df0));
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(og[u]) : "f"(dg1), "f"(dg0));
            }
          }
          unsigned char* os = ds + 12288 + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t off = (uint32_t)((sub * 4 + j) ^ (lane & 7)) << 4;
            *reinterpret_cast<uint4*>(os + off) = qf[j];
            *reinterpret_cast<uint4*>(os + 4096 + off) = qg[j];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          release_acc(acc);
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&mapC16), "r"(smem_u32(ds + 12288)), "r"(cf), "r"(row0) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&mapC16), "r"(smem_u32(ds + 16384)), "r"(128 + cf), "r"(row0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (p.colsum) {
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c = lane + 32 * cc;
              float sacc = 0.f;
#pragma unroll
              for (int rr = 0; rr < 32; ++rr)
                sacc += __half2float(*reinterpret_cast<const __half*>(ds + 12288 + b * 4096 + rr * 128 + ((uint32_t)((c >> 3) ^ (rr & 7)) << 4) + (c & 7) * 2));
              atomicAdd(&bias_s[128 * b + cf + c], sacc);
            }
        }
        continue;
      }
      if (BN == 256 && p.gate) {
        // ---- gated activation in the epilogue (block_wide16.cu): the tile is [f 128 | g 128]; a warp takes 64 f columns and
        // the matching 64 g columns of its 32 rows, stores both (the saved pre-activations) and z = tanh(f) sigmoid(g) ----
        unsigned char* gs = out_stage + ew * 3 * 4096;        // blocks: f, g, z ([32 rows][64 halfs], 128B swizzle)
        const int cf = (ew >> 2) * 64, cg = 128 + cf;
        mbar_wait(&tmem_full_bar[acc], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t va[32], vb[32];
          tld(va, cf + 32 * sub);
          tld(vb, cg + 32 * sub);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float f[32], g[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { f[j] = __uint_as_float(va[j]); g[j] = __uint_as_float(vb[j]); }
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bf = *reinterpret_cast<const float4*>(bias_s + cf + 32 * sub + j);
              const float4 bg = *reinterpret_cast<const float4*>(bias_s + cg + 32 * sub + j);
              f[j] += bf.x; f[j + 1] += bf.y; f[j + 2] += bf.z; f[j + 3] += bf.w;
              g[j] += bg.x; g[j + 1] += bg.y; g[j + 2] += bg.z; g[j + 3] += bg.w;
            }
          }
          uint4 hq[4];
          uint32_t* hw = reinterpret_cast<uint32_t*>(hq);
          unsigned char* os = gs + lane * 128;
#pragma unroll
          for (int j = 0; j < 16; ++j) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw[j]) : "f"(f[2 * j + 1]), "f"(f[2 * j]));
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(os + ((uint32_t)((sub * 4 + j) ^ (lane & 7)) << 4)) = hq[j];
#pragma unroll
          for (int j = 0; j < 16; ++j) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw[j]) : "f"(g[2 * j + 1]), "f"(g[2 * j]));
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(os + 4096 + ((uint32_t)((sub * 4 + j) ^ (lane & 7)) << 4)) = hq[j];
#pragma unroll
          for (int j = 0; j < 32; ++j) {      // two MUFU.TANH per element: sigmoid(g) = 0.5 tanh(g / 2) + 0.5  (2^-11 relative, the
            float t, u;                       // precision z is stored with; ex2 / rcp forms measured: +6 us per layer, same logits error)
            asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(f[j]));
            asm("tanh.approx.f32 %0, %1;" : "=f"(u) : "f"(0.5f * g[j]));
            f[j] = t * fmaf(0.5f, u, 0.5f);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw[j]) : "f"(f[2 * j + 1]), "f"(f[2 * j]));
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(os + 8192 + ((uint32_t)((sub * 4 + j) ^ (lane & 7)) << 4)) = hq[j];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          release_acc(acc);
          if (p.has_c16) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&mapC16), "r"(smem_u32(gs)), "r"(n0 + cf), "r"(row0) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&mapC16), "r"(smem_u32(gs + 4096)), "r"(n0 + cg), "r"(row0) : "memory");
          }
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&mapC), "r"(smem_u32(gs + 8192)), "r"(cf), "r"(row0) : "memory");      // z: mapC is the layer's column block of Zcat16
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        continue;
      }
      if (p.wstage) {
        // ---- wide staging (see UmmaParams::wstage): per tile ONE staging pass, one fence, one store group per warp ----
        constexpr int NBW = BN / 128;                         // 64-column blocks per warp (its BN / 2 columns)
        constexpr uint32_t WS = NBW * 4096;                   // bytes per warp
        unsigned char* ws_out = out_stage + ew * WS;
        unsigned char* ws_aux = out_stage + 8 * WS + ew * WS;
        if (p.aux && lane == 0) {                             // fp16 addend blocks of this warp's columns: in flight while the MMAs run
          int nblk = 0;
#pragma unroll
          for (int b = 0; b < NBW; ++b) nblk += (n0 + col_lo + 64 * b < p.N) ? 1 : 0;
          if (nblk) {
            mbar_expect_tx(&aux_bar[ew][0], nblk * 4096);
#pragma unroll
            for (int b = 0; b < NBW; ++b)
              if (n0 + col_lo + 64 * b < p.N) tma_load_2d(ws_aux + b * 4096, &mapAux, &aux_bar[ew][0], n0 + col_lo + 64 * b, row0);
          }
        }
        mbar_wait(&tmem_full_bar[acc], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous tile's stores have left the blocks
        __syncwarp();
        if (p.aux && n0 + col_lo < p.N) { mbar_wait(&aux_bar[ew][0], aux_cnt & 1u); ++aux_cnt; }
        auto chunk = [&](uint32_t (&v)[32], int c0, int b, int sub) {      // 32 columns [c0, c0 + 32): block b, half sub
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          const int nb = n0 + c0;
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bias_s + c0 + j);
              f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
            }
          }
          if (p.flags & GEMM_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            if (p.mask_out && row_ok && nb < p.N) {
              uint32_t word = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) word |= (f[j] > 0.f ? 1u : 0u) << j;
              p.mask_out[(size_t)row * p.ldmw + (nb >> 5)] = word;
            }
          }
          if (p.mask_in) {
            uint32_t word = mw[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) word = (c0 >> 5) == q ? mw[q] : word;
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = ((word >> j) & 1u) ? f[j] : 0.f;
          }
          if (p.aux) {
            const unsigned char* as = ws_aux + b * 4096 + lane * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 q = *reinterpret_cast<const uint4*>(as + ((uint32_t)((sub * 4 + j) ^ (lane & 7)) << 4));
              const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 a = __half22float2(h2[u]);
                f[8 * j + 2 * u] = fmaf(p.aux_scale, a.x, f[8 * j + 2 * u]);
                f[8 * j + 2 * u + 1] = fmaf(p.aux_scale, a.y, f[8 * j + 2 * u + 1]);
              }
            }
          }
          uint4 hq[4];
          uint32_t* hw = reinterpret_cast<uint32_t*>(hq);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw[j]) : "f"(f[2 * j + 1]), "f"(f[2 * j]));
          unsigned char* os = ws_out + b * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(os + ((uint32_t)((sub * 4 + j) ^ (lane & 7)) << 4)) = hq[j];
        };
#pragma unroll 1
        for (int b = 0; b < NBW; ++b) {
          const int c0 = col_lo + 64 * b;
          if (n0 + c0 >= p.N) break;                          // warp-uniform
          uint32_t va[32], vb[32];
          tld(va, c0);
          tld(vb, c0 + 32);                                   // (columns past N: the accumulator holds zeros / stale data, the store clips them)
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          chunk(va, c0, b, 0);
          chunk(vb, c0 + 32, b, 1);
        }
        // the accumulator is free as soon as it has been read: the MMAs of the tile after next start while the stores drain
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          release_acc(acc);
#pragma unroll
          for (int b = 0; b < NBW; ++b)
            if (n0 + col_lo + 64 * b < p.N)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&mapC16), "r"(smem_u32(ws_out + b * 4096)), "r"(n0 + col_lo + 64 * b), "r"(row0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (p.colsum) {      // lane: columns lane, lane + 32 of every 64-column block over the warp's 32 rows (rows past M hold zeros)
#pragma unroll 1
          for (int b = 0; b < NBW; ++b) {
            if (n0 + col_lo + 64 * b >= p.N) break;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c = lane + 32 * cc;
              float sacc = 0.f;
#pragma unroll
              for (int rr = 0; rr < 32; ++rr)
                sacc += __half2float(*reinterpret_cast<const __half*>(ws_out + b * 4096 + rr * 128 + ((uint32_t)((c >> 3) ^ (rr & 7)) << 4) + (c & 7) * 2));
              atomicAdd(&bias_s[col_lo + 64 * b + c], sacc);
            }
          }
        }
        continue;
      }
      mbar_wait(&tmem_full_bar[acc], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int hc = 0; hc < (p.m2 ? 2 : 1); ++hc)      // 256-row tiles: rows 128.. live in the second accumulator
#pragma unroll 1
      for (int c0 = col_lo; c0 < col_hi; c0 += 32) {
        if (n0 + c0 >= p.N) break;                      // warp-uniform
        uint32_t v[32];
        {
          const uint32_t taddr = tmem + (acc + hc) * BN + ((uint32_t)(quad * 32) << 16) + c0;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        const int nb = n0 + c0;
        const bool full = (nb + 32 <= p.N);
        if (p.flags & GEMM_ATOMIC) {
          if (p.c_scale != 0.f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * p.c_scale);
          }
          if (row + hc * UM < p.M) {
            float* dst = p.C + (size_t)(row + hc * UM) * p.ldc + nb;
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
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bias_s + c0 + j);
            f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
          }
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
          if (p.mask_out && row_ok) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) word |= (f[j] > 0.f ? 1u : 0u) << j;
            p.mask_out[(size_t)row * p.ldmw + (nb >> 5)] = word;
          }
        }
        if (p.mask_in) {
          uint32_t word = mw[0];
#pragma unroll
          for (int q = 1; q < 8; ++q) word = (c0 >> 5) == q ? mw[q] : word;
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = ((word >> j) & 1u) ? f[j] : 0.f;
        }
        if (p.aux) {
          const uint32_t buf = aux_cnt & 1u;
          if (lane == 0 && c0 + 32 < col_hi && nb + 32 < p.N) issue_aux(c0 + 32);   // next chunk, other buffer
          mbar_wait(&aux_bar[ew][buf], (aux_cnt >> 1) & 1u);
          const unsigned char* ms = my_aux + buf * STG_BYTES + lane * 128;
          if (p.aux_add) {      // [32 rows][32 halfs], 64-byte rows, 64B swizzle (as the fp16 output staging blocks)
            const unsigned char* hs = my_aux + buf * STG_BYTES + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 q = *reinterpret_cast<const uint4*>(hs + ((uint32_t)(j ^ ((lane >> 1) & 3)) << 4));
              const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 a = __half22float2(h2[u]);
                f[8 * j + 2 * u] = fmaf(p.aux_scale, a.x, f[8 * j + 2 * u]);
                f[8 * j + 2 * u + 1] = fmaf(p.aux_scale, a.y, f[8 * j + 2 * u + 1]);
              }
            }
          } else
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 mk = *reinterpret_cast<const float4*>(ms + ((uint32_t)(j ^ (lane & 7)) << 4));
            f[4 * j] = mk.x > 0.f ? f[4 * j] : 0.f;
            f[4 * j + 1] = mk.y > 0.f ? f[4 * j + 1] : 0.f;
            f[4 * j + 2] = mk.z > 0.f ? f[4 * j + 2] : 0.f;
            f[4 * j + 3] = mk.w > 0.f ? f[4 * j + 3] : 0.f;
          }
          __syncwarp();      // every lane has read the buffer before a later load re-targets it
          ++aux_cnt;
        }
        uint4 hq[4];      // fp16 copy of the (unscaled) values
        if (p.has_c16) {
          uint32_t* hw = reinterpret_cast<uint32_t*>(hq);      // (one saturating conversion per pair: +-65504 instead of inf)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hw[j]) : "f"(f[2 * j + 1]), "f"(f[2 * j]));
        }
        // (scaling and tf32 rounding only concern the fp32 output)
        if ((p.C || p.CT) && p.c_scale != 0.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= p.c_scale;
        }
        if ((p.C || p.CT) && (p.flags & GEMM_ROUND)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = round_tf32(f[j]);
        }
        if (staged) {   // row chunk -> swizzled staging block -> one TMA store (clipped to [M, N] by the tensor map)
          // (measured: a TMA store takes ~1500 cycles to have READ its staging block -- with two blocks per warp
          // that wait, not the arithmetic, set the pace of every GEMM with a short K loop)
          const uint32_t buf = out_cnt % (uint32_t)p.nbuf;
          if (lane == 0) {      // the store issued nbuf chunks ago has left its staging block
            if (p.nbuf == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            else if (p.nbuf == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          __syncwarp();
          unsigned char* os = my_out + buf * STG_BYTES + lane * 128;
          if (p.C) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(os + ((uint32_t)(j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
          if (p.has_c16) {      // [32 rows][32 halfs] (64-byte rows, 64B swizzle) for the next GEMM
            unsigned char* hs = my_c16 + buf * (STG_BYTES / 2) + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(hs + ((uint32_t)(j ^ ((lane >> 1) & 3)) << 4)) = hq[j];
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (p.C)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&mapC), "r"(smem_u32(my_out + buf * STG_BYTES)), "r"(nb), "r"(row0) : "memory");
            if (p.has_c16)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&mapC16), "r"(smem_u32(my_c16 + buf * (STG_BYTES / 2))), "r"(nb), "r"(row0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (p.colsum) {      // lane j: column j of the fp16 block over its 32 rows (rows past M and columns past N hold zeros)
            const unsigned char* hb = my_c16 + buf * (STG_BYTES / 2);
            float sacc = 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              sacc += __half2float(*reinterpret_cast<const __half*>(hb + rr * 64 + ((uint32_t)((lane >> 3) ^ ((rr >> 1) & 3)) << 4) + (lane & 7) * 2));
            atomicAdd(&bias_s[c0 + lane], sacc);
          }
          ++out_cnt;
        }
        if (p.CT && row_ok) {   // transposed copy: for a fixed column the 32 lanes write 32 consecutive floats
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || nb + j < p.N) p.CT[(size_t)(nb + j) * p.ldct + row] = f[j];
        }
      }
      // this warp has read everything it needs from the accumulator: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) release_acc(acc);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output stores have landed
    if (p.colsum) colsum_flush(-1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) {      // neither CTA leaves (or frees its TMEM) while the other may still touch its shared memory / barriers / TMEM
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * BN));
  } else {
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * BN));
  }
}

// shared memory of the epilogue staging regions, and the deepest output ring that fits
static size_t staging_bytes(bool has_c, bool has_aux, bool has_c16, int nbuf, int ewarps) {
  return (has_c ? ewarps * (size_t)nbuf * STG_BYTES : 0) + (has_aux ? ewarps * 2 * (size_t)STG_BYTES : 0) +
         (has_c16 ? ewarps * (size_t)nbuf * (STG_BYTES / 2) : 0);
}
constexpr size_t SMEM_OPTIN = 232448 - 2048;      // 227 KB per CTA minus the kernel's static shared memory (barriers, bias row)
// 8 epilogue warps when their staging fits next to the operand pipeline, else 4; two staging blocks per warp (deeper
// rings were measured: no gain)
static void pick_epilogue(size_t pipeline_bytes, bool has_c, bool has_aux, bool has_c16, UmmaParams& p) {
  p.nbuf = 2;
  p.ewarps = (pipeline_bytes + staging_bytes(has_c, has_aux, has_c16, 2, 8) <= SMEM_OPTIN) ? 8 : 4;
}

template <int BN, int PAIR>
static int launch_umma_t(dim3 grid, size_t smem, cudaStream_t st, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mC,
                         const CUtensorMap& mAux, const CUtensorMap& mC16, const UmmaParams& p) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gemm_umma_kernel<BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_OPTIN);
    attr = true;
  }
  if (!PAIR) {
    gemm_umma_kernel<BN, PAIR><<<grid, THREADS, smem, st>>>(mA, mB, mC, mAux, mC16, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_umma_kernel<BN, PAIR>, mA, mB, mC, mAux, mC16, p);
    if (e != cudaSuccess) return (int)e;
  }
  WN_CHECK_LAUNCH();
  return 0;
}
static int launch_umma(int BN, dim3 grid, size_t smem, cudaStream_t st, const CUtensorMap& mA, const CUtensorMap& mB,
                       const CUtensorMap& mC, const CUtensorMap& mAux, const CUtensorMap& mC16, const UmmaParams& p, bool pair = false) {
  if (smem > SMEM_OPTIN) return -5;
  if (pair) return BN == 256 ? launch_umma_t<256, 1>(grid, smem, st, mA, mB, mC, mAux, mC16, p) : -1;
  return BN == 256 ? launch_umma_t<256, 0>(grid, smem, st, mA, mB, mC, mAux, mC16, p)
                   : launch_umma_t<128, 0>(grid, smem, st, mA, mB, mC, mAux, mC16, p);
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
  CUtensorMap mC = mA, mAux = mA;   // (unused maps must still be valid objects)
  if (g.C && !(g.flags & GEMM_ATOMIC)) {
    rc = make_map(&mC, g.C, g.M, g.N, g.ldc, 32);
    if (rc) return rc;
  }
  if (g.aux) {
    rc = make_map(&mAux, g.aux, g.M, g.N, g.ldaux, 32);
    if (rc) return rc;
  }
  UmmaParams p;
  p.colsum = nullptr; p.colsum_scale = 0.f; p.a_split = 0; p.a_shift = 0; p.aux_add = 0; p.aux_scale = 0.f; p.tn_R = 0; p.tn_shift = 0; p.wstage = 0; p.gate = 0; p.dpre = 0;
  p.C = g.C; p.ldc = g.ldc; p.CT = CT; p.ldct = ldct; p.C2 = g.C2; p.ldc2 = g.ldc2; p.bias = g.bias;
  p.aux = g.aux; p.ldaux = g.ldaux; p.M = g.M; p.N = g.N; p.K = g.K; p.flags = g.flags;
  p.m2 = 0;
  p.a_mn = a_mn; p.b_mn = b_mn; p.f16 = 0; p.has_c16 = 0; p.mask_out = nullptr; p.mask_in = nullptr; p.ldmw = 0;
  int splits = (g.flags & GEMM_ATOMIC) ? (split_k > 0 ? split_k : 1) : 1;
  int kps = ((g.K + splits - 1) / splits + UK - 1) / UK * UK;
  splits = (g.K + kps - 1) / kps;
  p.k_per_split = kps;
  p.splits = splits;
  const int64_t items = (int64_t)((g.N + BN - 1) / BN) * ((g.M + UM - 1) / UM) * splits;
  if (items > (1 << 30)) return -1;
  dim3 grid((unsigned)(items < sm_count() ? items : sm_count()));
  const bool atomic = (g.flags & GEMM_ATOMIC) != 0;
  p.stages = atomic ? 4 : 3;
  const size_t pipe = 1024 + (size_t)p.stages * (UM * UK * 4 + BN * UK * 4);
  pick_epilogue(pipe, !atomic && g.C != nullptr, !atomic && g.aux != nullptr, false, p);
  const size_t smem = pipe + (atomic ? 0 : staging_bytes(g.C != nullptr, g.aux != nullptr, false, p.nbuf, p.ewarps));
  p.c_scale = 0.f;
  return launch_umma(BN, grid, smem, st, mA, mB, mC, mAux, mAux, p);
}

// fp16 operands (K-major), fp32 accumulate:  C[M,N] = c_scale * mask(relu(A16[M,K] . B16[N,K]^T + bias)), written as fp32
// and, when c16 is given, also as fp16 WITHOUT c_scale (the A operand of the next GEMM of a chain).
// Forward chain (skip sum, postprocess1/2): every GEMM of the step is bound by the L2 -> SM operand traffic, so halving
// the operand bytes (and doubling the MMA rate) nearly halves these.  Backward input-gradient chain: the gradients
// travel as fp16 in a domain scaled by a power of two (c_scale undoes it for the fp32 copies).
int gemm_f16_nt(const void* A16, int lda, const void* B16, int ldb, float* C, int ldc, void* C16, int ldc16, int M, int N,
                int K, const float* bias, const float* aux, int ldaux, float c_scale, int flags, cudaStream_t st,
                uint32_t* mask_out, const uint32_t* mask_in, int ldmw, float* colsum, float colsum_scale, const F16Extra* ex) {
  if (M <= 0 || N <= 0 || K <= 0 || !A16 || !B16 || (!C && !C16)) return -1;
  if (colsum && (!C16 || bias)) return -1;
  const int a_split = ex ? ex->a_split : 0;
  const void* aux16 = ex ? ex->aux16 : nullptr;
  if (a_split && (K != 2 * a_split || (a_split & 63))) return -1;
  if (aux16 && (aux || (ex->ldaux16 & 7) || ((uintptr_t)aux16 & 15))) return -3;
  if ((lda & 7) || (ldb & 7) || (C && (ldc & 3)) || (C16 && (ldc16 & 7)) || ((uintptr_t)A16 & 15) || ((uintptr_t)B16 & 15) ||
      ((uintptr_t)C & 15) || ((uintptr_t)C16 & 15) || (aux && ((ldaux & 3) || ((uintptr_t)aux & 15))) || (flags & GEMM_ATOMIC))
    return -3;
  const int BN = N > 128 ? 256 : 128;
  CUtensorMap mA, mB, mC, mAux, mC16;
  int rc = make_map16(&mA, A16, M, a_split ? a_split : K, lda, UM, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  // CTA pairs (gemm_umma_kernel<BN, 1>): 256-row tiles, each CTA loads half of the B tile (WN_GEMM_PAIR=0: off)
  static const bool pair_env = [] { const char* e = getenv("WN_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  const bool pair = pair_env && BN == 256 && M >= 4 * UM && sm_count() >= 2;
  rc = make_map16(&mB, B16, N, K, ldb, pair ? BN / 2 : BN, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  mC = mA;
  if (C) {
    rc = make_map(&mC, C, M, N, ldc, 32);
    if (rc) return rc;
  }
  mAux = mA;
  if (aux) {
    rc = make_map(&mAux, aux, M, N, ldaux, 32);
    if (rc) return rc;
  }
  // wide staging (UmmaParams::wstage): fp16-only outputs; needs room for 8 warps x BN/2 columns (x 2 with an fp16 addend)
  static const bool ws_env = [] { const char* e = getenv("WN_GEMM_WSTAGE"); return !(e && e[0] == '0'); }();
  const size_t ws_bytes = (size_t)8 * (BN / 128) * 4096 * (aux16 ? 2 : 1);
  const size_t stage_bytes = UM * UK * 4 + BN * UK * 4 / (pair ? 2 : 1);
  const int max_stages = pair ? MAX_STAGES : 4;
  int ws_stages = 0;
  if (ws_env && !C && C16 && !aux)
    for (int sg = max_stages; sg >= 2 && !ws_stages; --sg)
      if (1024 + sg * stage_bytes + ws_bytes <= SMEM_OPTIN) ws_stages = sg;
  if (ws_stages < 3 && K > 128) ws_stages = 0;      // (two stages only for the shortest K loops)
  // Measured: the wide staging wins where the epilogue sets the pace (K <= 256, relu-gradient masks, column sums, fp16
  // addends: post2_dgrad 80 -> 59 us, post1_dgrad 88 -> 79) and loses where it costs the operand ring its fourth stage
  // on a product that is bound by the L2 -> SM operand traffic (skip_fwd 176 -> 186, skip_dgrad 203 -> 213)
  {
    const bool has_aux_ = aux != nullptr || aux16 != nullptr;
    const int base_stages = (1024 + 4 * stage_bytes + staging_bytes(C != nullptr, has_aux_, C16 != nullptr, 2, 8) <= SMEM_OPTIN) ? 4 : 3;
    if (ws_stages < base_stages && K > 256 && !mask_in && !colsum) ws_stages = 0;      // (never with pairs: their 32 KB stages leave room)
  }
  const void* dpre_P16 = ex ? ex->dpre_P16 : nullptr;
  if (dpre_P16 && (N != 128 || C || aux || !aux16 || !C16 || bias || a_split || (ex->ldp & 7) || ((uintptr_t)dpre_P16 & 15))) return -3;
  if (dpre_P16) {
    ws_stages = 0;
    rc = make_map16(&mC, dpre_P16, M, 256, ex->ldp, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  void* gate_z16 = ex ? ex->gate_z16 : nullptr;
  if (gate_z16 && (N != 256 || C || aux || aux16 || !C16 || (ex->ldz & 7) || ((uintptr_t)gate_z16 & 15))) return -3;
  if (gate_z16) ws_stages = 0;
  const bool wstage = ws_stages > 0;
  if (gate_z16) {      // z block map in the mapC slot: [M][128] halfs at the layer's columns of Zcat16
    rc = make_map16(&mC, gate_z16, M, N / 2, ex->ldz, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (aux16) {
    rc = (wstage || dpre_P16) ? make_map16(&mAux, aux16, M, N, ex->ldaux16, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map16(&mAux, aux16, M, N, ex->ldaux16, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  mC16 = mA;
  if (C16) {
    rc = (wstage || gate_z16 || dpre_P16) ? make_map16(&mC16, C16, M, dpre_P16 ? 256 : N, ldc16, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map16(&mC16, C16, M, N, ldc16, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  UmmaParams p;
  p.colsum = nullptr; p.colsum_scale = 0.f;
  p.C = C; p.ldc = ldc; p.CT = nullptr; p.ldct = 0; p.C2 = nullptr; p.ldc2 = 0; p.bias = bias;
  p.aux = aux16 ? (const float*)aux16 : aux; p.ldaux = ldaux; p.M = M; p.N = N; p.K = K; p.flags = flags;
  p.m2 = 0;
  p.tn_R = 0; p.tn_shift = 0;
  p.a_split = a_split; p.a_shift = ex ? ex->a_shift : 0; p.aux_add = aux16 ? 1 : 0; p.aux_scale = ex ? ex->aux_scale : 1.f;
  p.a_mn = 0; p.b_mn = 0; p.f16 = 1; p.has_c16 = C16 ? 1 : 0;
  p.mask_out = mask_out; p.mask_in = mask_in; p.ldmw = ldmw;
  p.c_scale = (c_scale == 1.f) ? 0.f : c_scale;
  p.colsum = colsum; p.colsum_scale = colsum_scale;
  p.k_per_split = (K + 63) / 64 * 64;
  p.splits = 1;
  p.stages = 3;
  const int tm_rows = pair ? 2 * UM : UM;
  const int64_t items = (int64_t)((N + BN - 1) / BN) * ((M + tm_rows - 1) / tm_rows);
  unsigned gx = (unsigned)(items < sm_count() ? items : sm_count());
  if (pair) { gx = (unsigned)(2 * items < sm_count() ? 2 * items : sm_count()); gx &= ~1u; }      // whole pairs
  dim3 grid(gx);
  size_t pipe = 1024 + (size_t)p.stages * stage_bytes;
  // more operand stages where the (8-warp) staging leaves room for them: fp16-only outputs, CTA pairs
  const bool has_aux = aux != nullptr || aux16 != nullptr;
  while (p.stages < max_stages && pipe + stage_bytes + staging_bytes(C != nullptr, has_aux, C16 != nullptr, 2, 8) <= SMEM_OPTIN) {
    ++p.stages;
    pipe += stage_bytes;
  }
  pick_epilogue(pipe, C != nullptr, has_aux, C16 != nullptr, p);
  size_t smem = pipe + staging_bytes(C != nullptr, has_aux, C16 != nullptr, p.nbuf, p.ewarps);
  p.wstage = wstage ? 1 : 0;
  if (wstage) {
    p.stages = ws_stages;
    p.ewarps = 8;
    smem = 1024 + ws_stages * stage_bytes + ws_bytes;
  }
  p.dpre = dpre_P16 ? 1 : 0;
  if (dpre_P16) {      // five blocks per warp (dz_skip, f, g in; df, dg out) beside a two-stage ring (K = 128: two k-steps per tile)
    p.stages = 2;
    p.ewarps = 8;
    smem = 1024 + 2 * stage_bytes + (size_t)8 * 5 * 4096;
  }
  p.gate = gate_z16 ? 1 : 0;
  if (gate_z16) {      // three staging blocks per warp (f, g, z); as many operand stages as fit beside them
    const size_t gs_bytes = (size_t)8 * 3 * 4096;
    int sg = max_stages;
    while (sg > 2 && 1024 + sg * stage_bytes + gs_bytes > SMEM_OPTIN) --sg;
    p.stages = sg;
    p.ewarps = 8;
    smem = 1024 + sg * stage_bytes + gs_bytes;
  }
  return launch_umma(BN, grid, smem, st, mA, mB, mC, mAux, mC16, p, pair);
}

// fp16 [rows][cols] row-major viewed as [cols/64][rows][64]: one box {64, 64 rows, n_blocks} lands as n_blocks
// consecutive [64 rows][64 halfs] blocks (128-byte rows, plain 128B swizzle).  Coordinates: (0, row, col / 64).
static int make_map16_blocks_mn(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int n_blocks) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -8;
  cuuint64_t gdim[3] = {64, (cuuint64_t)rows, (cuuint64_t)(cols / 64)};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, 128};
  cuuint32_t box[3] = {64, 64, (cuuint32_t)n_blocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -9;
}

bool gemm_f16_tn_supported(int lda, int ldb, int M, int N) {
  return M > 0 && N > 0 && !(M & 63) && !(N & 63) && !(lda & 7) && !(ldb & 7);
}

// Weight-gradient form on fp16 operands:  C[M,N] += c_scale * A16[K,M]^T . B16[K,N]  (both operands as they lie in
// memory: the contracted dimension -- time -- is the row index), split over K, accumulated with red.global.add.
int gemm_f16_tn(const void* A16, int lda, const void* B16, int ldb, float* C, int ldc, int M, int N, int K, float c_scale,
                int split_k, cudaStream_t st, int tn_R, int tn_shift) {
  if (K <= 0 || !A16 || !B16 || !C) return -1;
  if (tn_R && (M != 2 * tn_R || (tn_R & 127) || tn_shift < 0)) return -1;
  if (!gemm_f16_tn_supported(lda, ldb, M, N) || ((uintptr_t)A16 & 15) || ((uintptr_t)B16 & 15) || (ldc & 3) || ((uintptr_t)C & 15))
    return -3;
  const int BN = N > 128 ? 256 : 128;
  // 256-row tiles (two MMAs per k-step on one B stage): the GEMM is bound by the L2 -> SM operand traffic, and a
  // 256 x 256 tile moves 64 KB per k-step where two 128 x 256 tiles move 96 KB (WN_GEMM_M2=0 switches it off)
  static int m2_on = -1;
  if (m2_on < 0) m2_on = (getenv("WN_GEMM_M2") && atoi(getenv("WN_GEMM_M2")) == 0) ? 0 : 1;
  const int m2 = (m2_on && M >= 1024 && BN == 256 && !tn_R) ? 1 : 0;      // (few output tiles -> many K splits -> the atomics of the bigger tiles cost more than the operand traffic saved: measured on the 512-row products)
  const int TMR = m2 ? 2 * UM : UM;
  // CTA pairs for the 256-column products that do not use the 256-row tiles: each CTA loads its 128 A columns and half of B.
  // Correct (tests pass with it on) but measured without gain -- post1_wgrad 71.7 us either way, post2_wgrad 55 vs 51, the
  // two-tap weight gradient of the wide blocks unchanged: these split-K products are paced by their red.add flushes, not by
  // the operand traffic -- so it is OFF unless WN_GEMM_PAIR_TN=1.
  static const bool pair_env = [] { const char* e = getenv("WN_GEMM_PAIR_TN"); return e && e[0] == '1'; }();
  const bool pair = pair_env && !m2 && BN == 256 && M >= 2 * UM && (M % (2 * UM)) == 0 && sm_count() >= 2;
  CUtensorMap mA, mB;
  int rc = make_map16_blocks_mn(&mA, A16, K, tn_R ? tn_R : M, lda, TMR / 64);
  if (rc) return rc;
  rc = make_map16_blocks_mn(&mB, B16, K, N, ldb, (pair ? BN / 2 : BN) / 64);
  if (rc) return rc;
  UmmaParams p;
  p.colsum = nullptr; p.colsum_scale = 0.f; p.a_split = 0; p.a_shift = 0; p.aux_add = 0; p.aux_scale = 0.f; p.tn_R = 0; p.tn_shift = 0; p.wstage = 0; p.gate = 0; p.dpre = 0;
  p.C = C; p.ldc = ldc; p.CT = nullptr; p.ldct = 0; p.C2 = nullptr; p.ldc2 = 0; p.bias = nullptr;
  p.aux = nullptr; p.ldaux = 0; p.M = M; p.N = N; p.K = K; p.flags = GEMM_ATOMIC;
  p.nbuf = 2; p.ewarps = 8; p.m2 = m2; p.tn_R = tn_R; p.tn_shift = tn_shift;
  p.a_mn = 1; p.b_mn = 1; p.f16 = 1; p.has_c16 = 0; p.mask_out = nullptr; p.mask_in = nullptr; p.ldmw = 0;
  p.c_scale = (c_scale == 1.f) ? 0.f : c_scale;
  int splits = split_k > 0 ? split_k : 1;
  if (m2) {      // about two work items per SM
    const int tiles = ((N + BN - 1) / BN) * ((M + TMR - 1) / TMR);
    splits = (2 * sm_count() + tiles - 1) / tiles;
    if (splits > K / 256) splits = K / 256 > 0 ? K / 256 : 1;
  }
  int kps = ((K + splits - 1) / splits + 63) / 64 * 64;
  splits = (K + kps - 1) / kps;
  p.k_per_split = kps;
  p.splits = splits;
  p.stages = m2 ? 3 : (pair ? 6 : 4);
  const int tile_rows = pair ? 2 * UM : TMR;
  const int64_t items = (int64_t)((N + BN - 1) / BN) * ((M + tile_rows - 1) / tile_rows) * splits;
  if (items > (1 << 30)) return -1;
  unsigned gx = (unsigned)(items < sm_count() ? items : sm_count());
  if (pair) { gx = (unsigned)(2 * items < sm_count() ? 2 * items : sm_count()); gx &= ~1u; }
  dim3 grid(gx);
  const size_t smem = 1024 + (size_t)p.stages * (TMR * UK * 4 + BN * UK * 4 / (pair ? 2 : 1));
  return launch_umma(BN, grid, smem, st, mA, mB, mA, mA, mA, p, pair);
}

// out16[n][k] = half(in[k][n])   (fp16 K-major weight copies for gemm_f16_nt; [K][N] fp32 row-major in)
__global__ void transpose_half_kernel(const float* __restrict__ in, int K, int N, __half* __restrict__ out, int ldo) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? in[(size_t)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) out[(size_t)n * ldo + k] = __float2half_rn(tile[threadIdx.x][i]);
  }
}
__global__ void to_half_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    __half2 h[2] = {__floats2half2_rn(v.x, v.y), __floats2half2_rn(v.z, v.w)};
    *reinterpret_cast<uint2*>(out + 4 * i) = *reinterpret_cast<uint2*>(h);
  }
}
int to_half(const float* in, void* out, int64_t n, cudaStream_t st) {
  if (n & 3) return -3;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  to_half_kernel<<<(int)blocks, 256, 0, st>>>(in, (__half*)out, n / 4);
  WN_CHECK_LAUNCH();
  return 0;
}
int transpose_half(const float* in, int K, int N, void* out, int ldo, cudaStream_t st) {
  dim3 grid((N + 31) / 32, (K + 31) / 32);
  transpose_half_kernel<<<grid, dim3(32, 8), 0, st>>>(in, K, N, (__half*)out, ldo);
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
