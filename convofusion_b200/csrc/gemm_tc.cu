// tcgen05 / TMEM / TMA GEMM for sm_100a:  y[M,N] (+)= act(A[M,K] . W[N,K]^T + bias), bf16 operands,
// fp32 accumulation in tensor memory.
//
// Both operands are K-major (activations row-major [M,K]; nn.Linear weights row-major [N,K]), the
// canonical "TN" case: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages 128 x 64 and BN x 64 bf16
// boxes into a STAGES-deep shared-memory ring; one elected thread issues tcgen05.mma
// (cta_group::1, kind::f16, M=128, N=BN, K=16) into a 128-lane x BN-column fp32 TMEM accumulator;
// four epilogue warps read it back with tcgen05.ld.32x32b (one accumulator row per thread) and
// apply bias / activation / residual-accumulate before storing.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Two CTAs fit per SM (<=113 KB smem,
// BN<=256 TMEM columns each) so one CTA's epilogue overlaps the other's main loop.
#include "common.cuh"
#include "rowvec.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <cstring>
#include <cstdlib>

namespace cfb {

namespace {

constexpr int BM = 128;      // UMMA M (cta_group::1)
constexpr int BK = 64;       // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 192;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// Multicast variants: the data (and the complete_tx on the barrier at the same offset) land in every CTA of `mask`.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// Bulk tensor store / reduce-add of one shared-memory box (async proxy); completion tracked by bulk groups.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- cta_group::2 (CTA pair) variants.  The pair shares one MMA: the leader (cluster rank 0) issues it, the operands
// are read from BOTH CTAs' shared memory at the same offsets, each CTA's tensor memory receives its own 128 rows.
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
// TMA load whose completion is signalled on a barrier that may live in the peer CTA (cluster-space address).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor layout):
// [0,14) addr>>4 | [16,30) LBO>>4 (=1, unused when swizzled) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
// | [46,48) version=1 | [61,64) layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#ifndef CFB_TC_TRACE
#define CFB_TC_TRACE 0
#endif
// Debug timeline of CTA (0,0,0): globaltimer ns at the phase boundaries (build with -DCFB_TC_TRACE=1).
__device__ unsigned long long g_tc_trace[16];
__device__ __forceinline__ void trace(int slot) {
  if constexpr (CFB_TC_TRACE != 0) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      g_tc_trace[slot] = t;
    }
  }
}

template <int BN, int STAGES>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int STATS_OFF = BAR_OFF + 128;                             // 128 rows x (sum, sumsq) floats
  static constexpr int TOTAL = BAR_OFF + 128 + 1024 + 1024;                   // barriers, row stats, alignment slack
};

// One 128 x BN output tile: rows [m0, min(m0+128, m_end)), columns [n0, n0+BN).
//
// Cluster multicast (CN x CM CTAs = CN consecutive n-tiles x CM consecutive m-tiles).  With 128x128 tiles a CTA pulls
// 32 KB from L2 per 64-deep K block for 1 M MACs (32 MAC/B), and the L2->SM path, not the tensor pipe, bounds the
// kernel.  The CN CTAs of a cluster row need the SAME A tile and the CM CTAs of a cluster column the SAME B tile, so
// each CTA loads only a 1/CN slice of A and a 1/CM slice of B and TMA-multicasts it into every sharer's shared
// memory (same offset, each sharer's own `full` barrier).  A stage may be overwritten only when every CTA that
// receives data from this producer has consumed it, so consumers release a stage with a multicast tcgen05.commit to
// the `empty` barrier of all CN + CM - 1 CTAs that feed them.
//
// LNC (LayerNorm over a cluster): the four n-tile CTAs of one 128-row block of a [M,512] residual update form a
// cluster.  Each CTA keeps its updated 128x128 sub-block in shared memory, publishes per-row (sum, sum of squares) of
// its 128 columns, and after one cluster barrier reads the other three CTAs' partials through distributed shared
// memory (mapa + ld.shared::cluster).  Every CTA then normalises its own sub-block and writes the next GEMM's bf16
// operand: the LayerNorm costs no global read and no kernel launch on the critical path.
template <int BN, int STAGES, int CN = 1, int CM = 1, bool LNC = false, bool TMA_ONLY = false>
__device__ __forceinline__ void gemm_tile(const CUtensorMap* tmA_p, const CUtensorMap* tmB_p, const CUtensorMap* tmO_p, const int m0,
                                          const int M /* first row NOT to store */, const int n0, const int N,
                                          const int K, const Epilogue& ep) {
  using S = Smem<BN, STAGES>;
  constexpr int CL = CN * CM;
  constexpr bool CLUSTERED = CL > 1 || LNC;
  static_assert(!LNC || (CL == 1 && BN == 128), "cluster LayerNorm: 4 x (128x128) tiles, no multicast");
  uint32_t crank = 0;
  if constexpr (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint32_t rn = crank % CN, rm = crank / CN;
  // CTAs sharing my A tile (same rm) and my B tile (same rn); their union feeds / is fed by me
  uint32_t mask_a = 0, mask_b = 0;
#pragma unroll
  for (int i = 0; i < CN; ++i) mask_a |= 1u << (rm * CN + i);
#pragma unroll
  for (int j = 0; j < CM; ++j) mask_b |= 1u << (rn + j * CN);
  const uint16_t mask_u = (uint16_t)(mask_a | mask_b);
  const CUtensorMap& tmA = *tmA_p;
  const CUtensorMap& tmB = *tmB_p;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base + S::BAR_OFF;           // STAGES barriers
  const uint32_t bar_empty = bar_full + STAGES * 8;      // STAGES barriers
  const uint32_t bar_acc = bar_empty + STAGES * 8;       // accumulator ready
  const uint32_t tmem_slot = bar_acc + 8;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + S::BAR_OFF + (2 * STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = K / BK;
  if (threadIdx.x == 0) trace(0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, CN + CM - 1);   // one release per CTA that consumes data I produce
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  if constexpr (CL > 1) {   // barriers of every CTA must be initialised before a peer multicasts into them
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  pdl_sync();   // everything above touched only shared/tensor memory; operands of the previous kernel are read below
  if (threadIdx.x == 0) trace(1);

  if (warp == 0) {
    if (lane == 0) {
      constexpr int A_SLICE = BM / CN, B_SLICE = BN / CM;   // rows this CTA loads for its sharers
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_empty + s * 8, ph ^ 1);
        const uint32_t sa = base + s * S::STAGE_BYTES, sb = sa + S::A_BYTES;
        mbar_expect_tx(bar_full + s * 8, S::STAGE_BYTES);
        if constexpr (CN > 1)
          tma_load_2d_mc(sa + rn * (A_SLICE * BK * 2), &tmA, kb * BK, m0 + (int)rn * A_SLICE, bar_full + s * 8, (uint16_t)mask_a);
        else
          tma_load_2d(sa, &tmA, kb * BK, m0, bar_full + s * 8);
        if constexpr (CM > 1)
          tma_load_2d_mc(sb + rm * (B_SLICE * BK * 2), &tmB, kb * BK, n0 + (int)rm * B_SLICE, bar_full + s * 8, (uint16_t)mask_b);
        else
          tma_load_2d(sb, &tmB, kb * BK, n0, bar_full + s * 8);
        if (kb == 0) trace(2);
      }
      trace(3);
    }
    if constexpr (LNC) {   // see warp 1
      __syncwarp();
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        tc_fence_after();
        if (kb == 0) trace(4);
        const uint32_t sa = base + s * S::STAGE_BYTES, sb = sa + S::A_BYTES;
        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advancing K inside the 128-byte swizzle row: +32 bytes = +2 in the (addr >> 4) field
          umma_bf16(tmem_acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        // frees the smem slot once these MMAs retire -- in my own CTA and in every CTA that feeds me
        if constexpr (CL > 1) umma_commit_mc(bar_empty + s * 8, mask_u);
        else umma_commit(bar_empty + s * 8);
      }
      umma_commit(bar_acc);
      trace(5);
    }
    if constexpr (LNC) {   // matches the epilogue warps' barrier between publishing and reading the row statistics
      __syncwarp();
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  } else {
    // ---- epilogue.  tcgen05.ld hands each thread one accumulator ROW, which is the wrong shape for global memory
    // (a warp would touch 32 different cache lines per instruction).  The pipeline stages are idle once bar_acc has
    // fired, so each warp transposes its 32 x BN slab through that shared memory and then walks its rows with the
    // lanes spread across the columns: every global access of the bias / residual / output is row-contiguous.
    const int q = warp & 3;
    if (TMA_ONLY || (!LNC && ep.tma_out)) {
      // ---- TMA epilogue.  Each warp owns a 32-row slab of the tile.  Per 32 accumulator columns it adds bias /
      // activation in registers, writes the values into a 128B-swizzled [32 rows x 128 B] box in the (now idle)
      // pipeline shared memory and lets one lane hand that box to the TMA unit: a plain tensor store, or an f32
      // reduce-add performed at the L2 for residual updates (out += v), so the SM never reads the residual and no
      // thread waits on a global load or store.  Rows past M are clipped by the tensor map.
      float* bias_s = reinterpret_cast<float*>(gen_base + S::STATS_OFF);
      const int te = threadIdx.x - 64;
      const bool row_bias = ep.bias && ep.bias_period != 1;   // bias[(row % period), :]: read per row below
      if (te < BN) bias_s[te] = (ep.bias && !row_bias) ? ep.bias[n0 + te] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      if (warp == 2 && lane == 0) trace(6);
      const int r0w = m0 + q * 32;
      if (r0w < M) {
        const uint32_t stg_w = base + (uint32_t)q * (BN / 32) * 4096u;
        const uint32_t sw = (uint32_t)(lane & 7);
        const bool out_bf = ep.out_bf16 != 0;
        const int act = ep.act;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j * 4);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
          if (row_bias) {
            const float* brow = ep.bias + (size_t)((r0w + lane) % ep.bias_period) * N + n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(brow + j * 4);
              v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
            }
          }
          if (act) {
            if (out_bf) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = act_apply_fast(v[j], act);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], act);
            }
          }
          if (!out_bf) {
            const uint32_t box = stg_w + (uint32_t)c * 4096u + (uint32_t)lane * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(box + (((uint32_t)j ^ sw) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (ep.accumulate) tma_reduce_add_2d(tmO_p, n0 + c * 32, r0w, stg_w + (uint32_t)c * 4096u);
              else tma_store_2d(tmO_p, n0 + c * 32, r0w, stg_w + (uint32_t)c * 4096u);
              tma_commit();
            }
          } else {
            const uint32_t box = stg_w + (uint32_t)(c >> 1) * 4096u + (uint32_t)lane * 128u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
                pk[e] = *reinterpret_cast<const uint32_t*>(&t);
              }
              st_shared_v4(box + (((uint32_t)((c & 1) * 4 + j) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
            if (c & 1) {
              fence_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(tmO_p, n0 + (c >> 1) * 64, r0w, stg_w + (uint32_t)(c >> 1) * 4096u);
                tma_commit();
              }
            }
          }
        }
        if (lane == 0) {
          if (ep.ln_out) {                 // a LayerNorm tail reads these rows back: the updates must have landed
            tma_wait_all0();
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
          } else {
            tma_wait_read0();              // the boxes must stay intact until the TMA unit has read them
          }
        }
      }
      if (warp == 2 && lane == 0) trace(7);
    } else if constexpr (!TMA_ONLY) {
    constexpr int PITCH = BN + 4;              // floats; +4 keeps the 128-bit transposed stores conflict-free
    constexpr int CPL = BN / 32;               // columns per lane: 4 (BN=128), 2, 1
    static_assert(4 * 32 * PITCH * 4 <= STAGES * S::STAGE_BYTES, "epilogue staging does not fit in the pipeline smem");
    float* stg = reinterpret_cast<float*>(gen_base) + q * 32 * PITCH;
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    if (warp == 2 && lane == 0) trace(6);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * PITCH + c * 32 + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
    if (warp == 2 && lane == 0) trace(7);
    const int n = n0 + lane * CPL;
    const int r0 = m0 + q * 32;
    float bias_v[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) bias_v[i] = (ep.bias && ep.bias_period == 1) ? ep.bias[n + i] : 0.f;
    const bool out_f32 = !ep.out_bf16;
    // Rows are processed RB at a time in three separate phases -- gather (shared-memory slab, residual, periodic
    // bias), compute, scatter -- so the RB global loads of a batch are all in flight together.  Interleaving load and
    // store per row serialises on the store->load ordering the compiler must assume (12 us instead of 1 us per tile).
    constexpr int RB = 8;
    const bool do_acc = out_f32 && ep.accumulate;
    const bool row_bias = ep.bias && ep.bias_period != 1;
#pragma unroll 1
    for (int rb = 0; rb < 32; rb += RB) {
      float v[RB][CPL], res[RB][CPL];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int rr = rb + i;
        if constexpr (CPL == 4) {
          const float4 t = *reinterpret_cast<const float4*>(stg + rr * PITCH + lane * 4);
          v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
        } else if constexpr (CPL == 2) {
          const float2 t = *reinterpret_cast<const float2*>(stg + rr * PITCH + lane * 2);
          v[i][0] = t.x; v[i][1] = t.y;
        } else {
          v[i][0] = stg[rr * PITCH + lane];
        }
      }
      if (do_acc) {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const int r = r0 + rb + i;
          if (r < M) {
            const float* o = reinterpret_cast<const float*>(ep.out) + (size_t)r * ep.ldo + n;
            if constexpr (CPL == 4) {
              const float4 t = *reinterpret_cast<const float4*>(o);
              res[i][0] = t.x; res[i][1] = t.y; res[i][2] = t.z; res[i][3] = t.w;
            } else if constexpr (CPL == 2) {
              const float2 t = *reinterpret_cast<const float2*>(o);
              res[i][0] = t.x; res[i][1] = t.y;
            } else {
              res[i][0] = *o;
            }
          }
        }
      }
      if (row_bias) {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const int r = r0 + rb + i;
          if (r < M) {
            const float* brow = ep.bias + (size_t)(r % ep.bias_period) * N + n;
#pragma unroll
            for (int c = 0; c < CPL; ++c) v[i][c] += brow[c];
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int c = 0; c < CPL; ++c) v[i][c] += bias_v[c];
      }
      if (ep.act) {
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int c = 0; c < CPL; ++c) v[i][c] = out_f32 ? act_apply(v[i][c], ep.act) : act_apply_fast(v[i][c], ep.act);
      }
      if (do_acc) {
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int c = 0; c < CPL; ++c) v[i][c] += res[i][c];
      }
      if constexpr (LNC) {   // keep the updated values on chip and publish this CTA's share of the row statistics
        float* stats = reinterpret_cast<float*>(gen_base + S::STATS_OFF);
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int c = 0; c < CPL; ++c) { s1 += v[i][c]; s2 = fmaf(v[i][c], v[i][c], s2); }
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          if (lane == 0) { stats[(q * 32 + rb + i) * 2] = s1; stats[(q * 32 + rb + i) * 2 + 1] = s2; }
          *reinterpret_cast<float4*>(stg + (rb + i) * PITCH + lane * 4) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        }
      }
      for (int rep = 0; rep < ep.replicate; ++rep) {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const int r = r0 + rb + i;
          if (r >= M) continue;
          const size_t off = (size_t)rep * ep.rep_stride + (size_t)r * ep.ldo + n;
          if (out_f32) {
            float* o = reinterpret_cast<float*>(ep.out) + off;
            if constexpr (CPL == 4) *reinterpret_cast<float4*>(o) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
            else if constexpr (CPL == 2) *reinterpret_cast<float2*>(o) = make_float2(v[i][0], v[i][1]);
            else *o = v[i][0];
          } else {
            bf16* o = reinterpret_cast<bf16*>(ep.out) + off;
            if constexpr (CPL == 4) {
              const __nv_bfloat162 t0 = __floats2bfloat162_rn(v[i][0], v[i][1]);
              const __nv_bfloat162 t1 = __floats2bfloat162_rn(v[i][2], v[i][3]);
              uint2 pk;
              pk.x = *reinterpret_cast<const uint32_t*>(&t0); pk.y = *reinterpret_cast<const uint32_t*>(&t1);
              *reinterpret_cast<uint2*>(o) = pk;
            } else if constexpr (CPL == 2) {
              *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(v[i][0], v[i][1]);
            } else {
              *o = __float2bfloat16_rn(v[i][0]);
            }
          }
        }
      }
    }
    if constexpr (LNC) {
      // ---- LayerNorm over the cluster: statistics of the other three column blocks through DSMEM
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      const uint32_t my_stats = base + S::STATS_OFF + (uint32_t)(q * 32 + lane) * 8;
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (uint32_t peer = 0; peer < 4; ++peer) {
        uint32_t ra;
        float a1, a2;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(my_stats), "r"(peer));
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(a1) : "r"(ra) : "memory");
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(a2) : "r"(ra + 4) : "memory");
        t1 += a1; t2 += a2;
      }
      const float mean_l = t1 * (1.0f / 512.0f);
      const float rstd_l = rsqrtf(fmaxf(t2 * (1.0f / 512.0f) - mean_l * mean_l, 0.f) + 1e-5f);   // row q*32+lane
      const float* mod = ep.ln_mod ? ep.ln_mod + (ep.ln_step ? (size_t)(*ep.ln_step) * ep.ln_mod_stride : 0) : nullptr;
      const float4 g4 = *reinterpret_cast<const float4*>(ep.ln_g + n);
      const float4 b4 = *reinterpret_cast<const float4*>(ep.ln_b + n);
      float4 sc4 = make_float4(0.f, 0.f, 0.f, 0.f), sh4 = sc4;
      if (mod) { sc4 = *reinterpret_cast<const float4*>(mod + n); sh4 = *reinterpret_cast<const float4*>(mod + 512 + n); }
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        const float mu = __shfl_sync(0xffffffffu, mean_l, rr), rs = __shfl_sync(0xffffffffu, rstd_l, rr);
        const float4 x = *reinterpret_cast<const float4*>(stg + rr * PITCH + lane * 4);
        float y0 = (x.x - mu) * rs * g4.x + b4.x, y1 = (x.y - mu) * rs * g4.y + b4.y;
        float y2 = (x.z - mu) * rs * g4.z + b4.z, y3 = (x.w - mu) * rs * g4.w + b4.w;
        if (mod) {
          y0 = act_apply(y0 * (1.0f + sc4.x) + sh4.x, CFB_ACT_SILU); y1 = act_apply(y1 * (1.0f + sc4.y) + sh4.y, CFB_ACT_SILU);
          y2 = act_apply(y2 * (1.0f + sc4.z) + sh4.z, CFB_ACT_SILU); y3 = act_apply(y3 * (1.0f + sc4.w) + sh4.w, CFB_ACT_SILU);
        }
        const int r = r0 + rr;
        if (r < M) {
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(y0, y1), p1 = __floats2bfloat162_rn(y2, y3);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&p0); pk.y = *reinterpret_cast<const uint32_t*>(&p1);
          *reinterpret_cast<uint2*>(ep.ln_out + (size_t)r * 512 + n) = pk;
        }
      }
    } else
    // ---- fused LayerNorm of this 128-row block by the CTA that completes it (see Epilogue::ln_out)
    if (ep.ln_out != nullptr) {
      __shared__ int s_last;
      __threadfence();                                            // my residual stores before my counter increment
      asm volatile("bar.sync 1, 128;" ::: "memory");              // the four epilogue warps only
      if (warp == 2 && lane == 0) {
        const int done = atomicAdd(ep.ln_counters + blockIdx.y, 1);
        s_last = (done == (int)gridDim.x - 1);
        if (s_last) ep.ln_counters[blockIdx.y] = 0;               // every n-tile of this block has arrived: re-arm
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (s_last) {
        __threadfence();                                          // order the counter observation before the reads
        const float* mod = ep.ln_mod ? ep.ln_mod + (ep.ln_step ? (size_t)(*ep.ln_step) * ep.ln_mod_stride : 0) : nullptr;
        constexpr int LD = 512, NR = 4;                           // rows in flight per warp
        float gam[16], bet[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 g4 = *reinterpret_cast<const float4*>(ep.ln_g + i * 128 + lane * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(ep.ln_b + i * 128 + lane * 4);
          gam[4 * i] = g4.x; gam[4 * i + 1] = g4.y; gam[4 * i + 2] = g4.z; gam[4 * i + 3] = g4.w;
          bet[4 * i] = b4.x; bet[4 * i + 1] = b4.y; bet[4 * i + 2] = b4.z; bet[4 * i + 3] = b4.w;
        }
        const float* hbase = reinterpret_cast<const float*>(ep.out);
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += NR) {
          float x[NR][16];
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            const int r = min(m0 + q * 32 + rb + j, M - 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 t = __ldcg(reinterpret_cast<const float4*>(hbase + (size_t)r * LD + i * 128 + lane * 4));
              x[j][4 * i] = t.x; x[j][4 * i + 1] = t.y; x[j][4 * i + 2] = t.z; x[j][4 * i + 3] = t.w;
            }
          }
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            const int r = m0 + q * 32 + rb + j;
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) sum += x[j][i];
            const float mu = warp_sum(sum) * (1.0f / LD);
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[j][i] -= mu; sq += x[j][i] * x[j][i]; }
            const float rstd = rsqrtf(warp_sum(sq) * (1.0f / LD) + 1e-5f);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[j][i] = x[j][i] * rstd * gam[i] + bet[i];
            if (mod) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 sc = *reinterpret_cast<const float4*>(mod + i * 128 + lane * 4);
                const float4 sh = *reinterpret_cast<const float4*>(mod + LD + i * 128 + lane * 4);
                x[j][4 * i] = act_apply(x[j][4 * i] * (1.0f + sc.x) + sh.x, CFB_ACT_SILU);
                x[j][4 * i + 1] = act_apply(x[j][4 * i + 1] * (1.0f + sc.y) + sh.y, CFB_ACT_SILU);
                x[j][4 * i + 2] = act_apply(x[j][4 * i + 2] * (1.0f + sc.z) + sh.z, CFB_ACT_SILU);
                x[j][4 * i + 3] = act_apply(x[j][4 * i + 3] * (1.0f + sc.w) + sh.w, CFB_ACT_SILU);
              }
            }
            if (r < M) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 t0 = __floats2bfloat162_rn(x[j][4 * i], x[j][4 * i + 1]);
                const __nv_bfloat162 t1 = __floats2bfloat162_rn(x[j][4 * i + 2], x[j][4 * i + 3]);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&t0); pk.y = *reinterpret_cast<const uint32_t*>(&t1);
                *reinterpret_cast<uint2*>(ep.ln_out + (size_t)r * LD + i * 128 + lane * 4) = pk;
              }
            }
          }
        }
      }
    }
    }   // staged (non-TMA) epilogue
  }
  if (warp == 2 && lane == 0) trace(8);
  tc_fence_before();
  if constexpr (CLUSTERED) {   // no CTA may exit while a peer can still multicast into it, arrive on its barriers or read its statistics
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (threadIdx.x == 0) trace(9);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, BN);
    if (lane == 0) trace(10);
  }
  if constexpr (TMA_ONLY) {
    // ---- LayerNorm tail (Epilogue::ln_out, float [M,512] residual output).  The gridDim.x CTAs of a 128-row block
    // signal a per-block counter once their updates are globally visible; each then waits for the whole block and
    // normalises its own share of the rows with all six warps, writing the next GEMM's bf16 operand.  CTAs are
    // dispatched in block-id order, so the CTAs a spinning CTA waits for are already resident or ahead of every
    // undispatched CTA: no deadlock.  The second counter re-arms both once every CTA of the block has passed.
    if (ep.ln_out != nullptr) {
      int* cnt = ep.ln_counters + 2 * blockIdx.y;
      const int n_cta = (int)gridDim.x;
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(cnt, 1);
        trace(11);
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_gpu(cnt) < n_cta) {
          __nanosleep(40);
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > 2000000ull) break;   // 2 ms: never hang the device on a protocol error
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) trace(12);
      const float* mod = ep.ln_mod ? ep.ln_mod + (ep.ln_step ? (size_t)(*ep.ln_step) * ep.ln_mod_stride : 0) : nullptr;
      const int per = (BM + n_cta - 1) / n_cta;
      const int r_lo = m0 + (int)blockIdx.x * per;
      const int r_hi = min(min(r_lo + per, m0 + BM), M);
      const float* hbase = reinterpret_cast<const float*>(ep.out);
      constexpr int NR = 3;                                   // rows in flight per warp
#pragma unroll 1
      for (int r = r_lo + warp * NR; r < r_hi; r += (NTHREADS / 32) * NR) {
        RowVec<512> rv[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) rv[j].load_cg(hbase + (size_t)min(r + j, r_hi - 1) * 512, lane);
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          ln_row_finish<512, true>(rv[j], ep.ln_g, ep.ln_b, mod, lane);
          if (r + j < r_hi) rv[j].store(ep.ln_out + (size_t)(r + j) * 512, lane);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        trace(13);
        const int old = atomicAdd(cnt + 1, 1);
        if (old == n_cta - 1) { cnt[1] = 0; __threadfence(); cnt[0] = 0; }
      }
    }
  }
}

template <int BN, int STAGES, int CN, int CM>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmO, int M, int N,
                                                           int K, Epilogue ep) {
  gemm_tile<BN, STAGES, CN, CM>(&tmA, &tmB, &tmO, blockIdx.y * BM, M, blockIdx.x * BN, N, K, ep);
}

// TMA-epilogue-only instantiation: no staged ld/st epilogue in the binary, so it fits 3 CTAs per SM (<=112 registers,
// 2-stage ring = 68 KB of shared memory): the phases of co-resident CTAs (prologue, operand latency, main loop,
// epilogue) overlap on one SM.
template <int BN, int STAGES, int OCC>
__global__ void __launch_bounds__(NTHREADS, OCC) gemm_tc_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmO, int M, int N,
                                                                  int K, Epilogue ep) {
  gemm_tile<BN, STAGES, 1, 1, false, true>(&tmA, &tmB, &tmO, blockIdx.y * BM, M, blockIdx.x * BN, N, K, ep);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tc_ln_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB, int M, int N,
                                                                  int K, Epilogue ep) {
  gemm_tile<BN, STAGES, 1, 1, true>(&tmA, &tmB, nullptr, blockIdx.y * BM, M, blockIdx.x * BN, N, K, ep);
}

// ---------------------------------------------------------------- cta_group::2: 256 x 256 tile per CTA pair
// With 128x128 tiles a CTA pulls 32 KB from L2 per 64-deep K block for 1 M MACs; at full occupancy the L2 -> SM path
// (not the tensor pipe) bounds the kernel (launch list at batch 256: 395 TFLOP/s).  A CTA pair on one TPC shares one
// tcgen05.mma.cta_group::2 (M = 256, N = 256): CTA r holds rows [m0 + 128 r, +128) of A and rows [n0 + 128 r, +128) of W
// -- the same 32 KB per K block as before -- and receives the accumulators of ITS 128 rows for all 256 columns in its
// own tensor memory, i.e. twice the MACs per byte pulled from L2.
//   * every barrier the MMA waits on lives in the leader (rank 0): both CTAs' TMA loads complete_tx on the leader's
//     `full` barrier, the leader's producer thread expects the bytes of both;
//   * a stage is released by ONE multicast tcgen05.commit that arrives on the `empty` barrier of both CTAs (the MMA
//     reads both CTAs' shared memory), and the accumulator-ready commit is multicast the same way;
//   * the epilogue is the TMA store / L2 reduce-add epilogue of gemm_tile, 256 columns in two passes of 128 so that the
//     staging boxes fit in the idle 96 KB pipeline ring.
template <int STAGES>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmO, int M, int N,
                                                                  int K, Epilogue ep) {
  using S = Smem<128, STAGES>;          // per CTA: 128 x 64 of A + 128 x 64 of W per stage
  constexpr int BN2 = 256;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool leader = crank == 0;
  const int m0 = (int)(blockIdx.x >> 1) * 256 + (int)crank * 128;
  const int n0 = (int)blockIdx.y * BN2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base + S::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_acc = bar_empty + STAGES * 8;
  const uint32_t tmem_slot = bar_acc + 8;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + S::BAR_OFF + (2 * STAGES + 1) * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = K / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);    // used in the leader only: its producer's arrive.expect_tx (+ both CTAs' bytes)
      mbar_init(bar_empty + s * 8, 1);   // one multicast commit per round
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, BN2);   // the same warp of both CTAs, same shared-memory slot
  tc_fence_before();
  cluster_sync_all();                                // peer barriers initialised, allocation visible
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_empty + s * 8, ph ^ 1);
        const uint32_t sa = base + s * S::STAGE_BYTES, sb = sa + S::A_BYTES;
        if (leader) mbar_expect_tx(bar_full + s * 8, 2 * S::STAGE_BYTES);
        const uint32_t full_leader = mapa_shared(bar_full + s * 8, 0);
        tma_load_2d_pair(sa, &tmA, kb * BK, m0, full_leader);
        tma_load_2d_pair(sb, &tmB, kb * BK, n0 + (int)crank * 128, full_leader);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(256, BN2);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        tc_fence_after();
        const uint32_t sa = base + s * S::STAGE_BYTES, sb = sa + S::A_BYTES;
        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16_pair(tmem_acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        umma_commit_pair(bar_empty + s * 8, (uint16_t)3);
      }
      umma_commit_pair(bar_acc, (uint16_t)3);
    }
  } else {
    const int q = warp & 3;
    float* bias_s = reinterpret_cast<float*>(gen_base + S::STATS_OFF);     // 256 floats
    const int te = threadIdx.x - 64;
    const bool row_bias = ep.bias && ep.bias_period != 1;
    bias_s[te] = (ep.bias && !row_bias) ? ep.bias[n0 + te] : 0.f;
    bias_s[te + 128] = (ep.bias && !row_bias) ? ep.bias[n0 + te + 128] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int r0w = m0 + q * 32;
    if (r0w < M) {
      const uint32_t stg_w = base + (uint32_t)q * 4u * 4096u;     // 4 boxes of [32 rows x 128 B] per warp and pass
      const uint32_t sw = (uint32_t)(lane & 7);
      const bool out_bf = ep.out_bf16 != 0;
      const int act = ep.act;
#pragma unroll 1
      for (int c = 0; c < BN2 / 32; ++c) {
        const int cb = c & 3;                                       // box slot inside the pass
        if (c == 4) {                                               // second pass reuses the boxes of the first
          if (lane == 0) tma_wait_read0();
          __syncwarp();
        }
        float v[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j * 4);
          v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
        }
        if (row_bias) {
          const float* brow = ep.bias + (size_t)((r0w + lane) % ep.bias_period) * N + n0 + c * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(brow + j * 4);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        }
        if (act) {
          if (out_bf) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = act_apply_fast(v[j], act);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], act);
          }
        }
        if (!out_bf) {
          const uint32_t box = stg_w + (uint32_t)cb * 4096u + (uint32_t)lane * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(box + (((uint32_t)j ^ sw) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                         __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (ep.accumulate) tma_reduce_add_2d(&tmO, n0 + c * 32, r0w, stg_w + (uint32_t)cb * 4096u);
            else tma_store_2d(&tmO, n0 + c * 32, r0w, stg_w + (uint32_t)cb * 4096u);
            tma_commit();
          }
        } else {
          const uint32_t box = stg_w + (uint32_t)(cb >> 1) * 4096u + (uint32_t)lane * 128u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
              pk[e] = *reinterpret_cast<const uint32_t*>(&t);
            }
            st_shared_v4(box + (((uint32_t)((c & 1) * 4 + j) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
          if (c & 1) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmO, n0 + (c >> 1) * 64, r0w, stg_w + (uint32_t)(cb >> 1) * 4096u);
              tma_commit();
            }
          }
        }
      }
      if (lane == 0) tma_wait_read0();
    }
  }
  tc_fence_before();
  cluster_sync_all();     // neither CTA may exit (or free tensor memory) while the pair's MMA / commits can still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_acc, BN2);
  }
}

// Grouped launch: blockIdx.z picks a group = (A map, B map, row block, bias, output).  All groups share N, K and
// the epilogue flags.  Used for the conditional rows of the guidance batch, where every branch owns one
// contiguous row block and its own stream's weights.
struct GroupedArgs {
  CUtensorMap tmA[TC_MAX_GROUPS];
  CUtensorMap tmB[TC_MAX_GROUPS];
  CUtensorMap tmO[TC_MAX_GROUPS];
  int row_start[TC_MAX_GROUPS];
  int rows[TC_MAX_GROUPS];
  const float* bias[TC_MAX_GROUPS];
  void* out[TC_MAX_GROUPS];
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tc_grouped_kernel(const __grid_constant__ GroupedArgs g, int N, int K,
                                                                   Epilogue ep) {
  const int z = blockIdx.z;
  const int m0 = g.row_start[z] + blockIdx.y * BM;
  const int m_end = g.row_start[z] + g.rows[z];
  if (m0 >= m_end) return;   // uniform for the whole CTA, before any barrier / TMEM allocation
  ep.bias = g.bias[z];
  ep.out = g.out[z];
  gemm_tile<BN, STAGES>(&g.tmA[z], &g.tmB[z], &g.tmO[z], m0, m_end, blockIdx.x * BN, N, K, ep);
}

template <int BN, int STAGES, int OCC>
__global__ void __launch_bounds__(NTHREADS, OCC) gemm_tc_grouped_tma_kernel(const __grid_constant__ GroupedArgs g, int N,
                                                                          int K, Epilogue ep) {
  const int z = blockIdx.z;
  const int m0 = g.row_start[z] + blockIdx.y * BM;
  const int m_end = g.row_start[z] + g.rows[z];
  if (m0 >= m_end) return;
  ep.bias = g.bias[z];
  ep.out = g.out[z];
  gemm_tile<BN, STAGES, 1, 1, false, true>(&g.tmA[z], &g.tmB[z], &g.tmO[z], m0, m_end, blockIdx.x * BN, N, K, ep);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

enum MapKind { MAP_OPERAND = 0, MAP_OUT_F32 = 1, MAP_OUT_BF16 = 2 };
struct MapKey {
  const void* p; int rows, cols, ld, box_rows, kind;
  bool operator==(const MapKey& o) const {
    return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && kind == o.kind;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.p);
    h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols;
    h = h * 1000003u ^ (size_t)k.ld;   h = h * 1000003u ^ (size_t)(k.box_rows * 4 + k.kind);
    return h;
  }
};

// Descriptors are pure functions of (pointer, shape), so they are cached for the process lifetime.
// MAP_OPERAND: bf16 [rows, cols] K-major operand, box 64 x box_rows.  MAP_OUT_*: output tile boxes of one warp's
// 32 rows x 128 bytes (32 floats / 64 bf16), same 128-byte swizzle.
int get_map(const void* p, int rows, int cols, int ld, int box_rows, CUtensorMap* out, int kind = MAP_OPERAND) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{p, rows, cols, ld, box_rows, kind};
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return CFB_OK; }
  EncodeTiledFn enc = get_encode();
  CFB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const int esz = kind == MAP_OUT_F32 ? 4 : 2;
  cuuint64_t gstr[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(kind == MAP_OUT_F32 ? 32 : BK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, kind == MAP_OUT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(p), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for ptr=%p rows=%d cols=%d ld=%d box=%d kind=%d", (int)r, p,
              rows, cols, ld, box_rows, kind);
    return CFB_ERR_CUDA;
  }
  cache.emplace(key, tm);
  *out = tm;
  return CFB_OK;
}

int g_tc_tma_epi = 1;   // env CFB_TC_TMA_EPI=0 keeps the shared-memory-staged ld/st epilogue
int g_tc_occ3 = 0;      // env CFB_TC_OCC3=0: 3-stage ring, 2 CTAs per SM instead of 2-stage ring, 3 CTAs per SM

// The TMA epilogue takes one output copy and (for bf16) tiles at least one 64-column box wide.
bool tma_epilogue_ok(const Epilogue& ep, int BN_) {
  const bool tail = ep.ln_out != nullptr && ep.ln_counters != nullptr && ep.ln_tail;
  return g_tc_tma_epi && ep.replicate == 1 && (ep.ln_out == nullptr || tail) &&
         (BN_ >= 64 || !ep.out_bf16);
}

template <int BN, int STAGES, int CN, int CM>
int launch(const bf16* A, int lda, const bf16* W, int ldw, int w_rows, int M, int N, int K, const Epilogue& ep_in,
           cudaStream_t st) {
  using S = Smem<BN, STAGES>;
  CUtensorMap ta, tb, to;
  Epilogue ep = ep_in;
  CFB_TRY(get_map(A, M, K, lda, BM / CN, &ta));          // each CTA loads (and multicasts) a 1/CN slice of the A tile
  CFB_TRY(get_map(W, w_rows, K, ldw, BN / CM, &tb));     // rows past w_rows read as zeros (TMA out-of-bounds fill)
  ep.tma_out = tma_epilogue_ok(ep, BN);
  if (ep.tma_out) CFB_TRY(get_map(ep.out, M, N, ep.ldo, 32, &to, ep.out_bf16 ? MAP_OUT_BF16 : MAP_OUT_F32));
  else to = ta;
  dim3 grid(ceil_div(N, BN), ceil_div(ceil_div(M, BM), CM) * CM);   // whole clusters; surplus m-tiles store nothing
  if constexpr (CN * CM == 1) {
    if (ep.tma_out && g_tc_occ3)
      launch_k(gemm_tc_tma_kernel<BN, 2, 3>, grid, NTHREADS, Smem<BN, 2>::TOTAL, st, ta, tb, to, M, N, K, ep);
    else if (ep.tma_out)
      launch_k(gemm_tc_tma_kernel<BN, STAGES, 2>, grid, NTHREADS, S::TOTAL, st, ta, tb, to, M, N, K, ep);
    else
      launch_k(gemm_tc_kernel<BN, STAGES, 1, 1>, grid, NTHREADS, S::TOTAL, st, ta, tb, to, M, N, K, ep);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CN; attr[0].val.clusterDim.y = CM; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CFB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, CN, CM>, ta, tb, to, M, N, K, ep));
  }
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int g_tc_pair = 0;            // env CFB_TC_2CTA: 1 = cta_group::2 256x256 tiles for N % 256 == 0
int g_tc_pair_min_rows = 0;   // env CFB_TC_2CTA_MIN_ROWS: only for GEMMs with at least this many rows

int launch_pair(const bf16* A, int lda, const bf16* W, int ldw, int w_rows, int M, int N, int K, const Epilogue& ep_in,
                cudaStream_t st) {
  using S = Smem<128, 3>;
  CUtensorMap ta, tb, to;
  Epilogue ep = ep_in;
  CFB_TRY(get_map(A, M, K, lda, 128, &ta));
  CFB_TRY(get_map(W, w_rows, K, ldw, 128, &tb));          // rows past w_rows read as zeros
  ep.tma_out = 1;
  CFB_TRY(get_map(ep.out, M, N, ep.ldo, 32, &to, ep.out_bf16 ? MAP_OUT_BF16 : MAP_OUT_F32));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * ceil_div(M, 256), N / 256); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 2;
  CFB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<3>, ta, tb, to, M, N, K, ep));
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

// [M,512] residual update + cluster LayerNorm: one 4-CTA cluster per 128-row block, PDL allowed.
int launch_ln(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep, cudaStream_t st) {
  using S = Smem<128, 3>;
  CUtensorMap ta, tb;
  CFB_TRY(get_map(A, M, K, lda, BM, &ta));
  CFB_TRY(get_map(W, N, K, ldw, 128, &tb));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(4, ceil_div(M, BM)); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 2;
  CFB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_ln_kernel<128, 3>, ta, tb, M, N, K, ep));
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

template <int BN, int STAGES>
int launch_grouped(const TcGroup* groups, int n_groups, int a_rows_total, int lda, int ldw, int N, int K,
                   const Epilogue& ep_in, cudaStream_t st) {
  using S = Smem<BN, STAGES>;
  GroupedArgs g;
  memset(&g, 0, sizeof(g));
  Epilogue ep = ep_in;
  ep.tma_out = tma_epilogue_ok(ep, BN);
  int max_rows = 0;
  for (int z = 0; z < n_groups; ++z) {
    CFB_TRY(get_map(groups[z].A, a_rows_total, K, lda, BM, &g.tmA[z]));
    CFB_TRY(get_map(groups[z].W, N, K, ldw, BN, &g.tmB[z]));
    if (ep.tma_out && groups[z].rows > 0)   // rows past the group's block are clipped by the map
      CFB_TRY(get_map(groups[z].out, groups[z].row_start + groups[z].rows, N, ep.ldo, 32, &g.tmO[z],
                      ep.out_bf16 ? MAP_OUT_BF16 : MAP_OUT_F32));
    g.row_start[z] = groups[z].row_start; g.rows[z] = groups[z].rows;
    g.bias[z] = groups[z].bias; g.out[z] = groups[z].out;
    if (groups[z].rows > max_rows) max_rows = groups[z].rows;
  }
  if (max_rows <= 0) return CFB_OK;
  dim3 grid(ceil_div(N, BN), ceil_div(max_rows, BM), n_groups);
  if (ep.tma_out && g_tc_occ3)
    launch_k(gemm_tc_grouped_tma_kernel<BN, 2, 3>, grid, NTHREADS, Smem<BN, 2>::TOTAL, st, g, N, K, ep);
  else if (ep.tma_out)
    launch_k(gemm_tc_grouped_tma_kernel<BN, STAGES, 2>, grid, NTHREADS, S::TOTAL, st, g, N, K, ep);
  else
    launch_k(gemm_tc_grouped_kernel<BN, STAGES>, grid, NTHREADS, S::TOTAL, st, g, N, K, ep);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

}  // namespace

int g_tc_cluster = 11;   // 10*CN_max + CM_max; 11 = no clusters

int tc_trace_read(unsigned long long out[16]) {
  CFB_CUDA(cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(unsigned long long) * 16));
  return CFB_OK;
}

int init_gemm_tc_kernels() {
  static bool done = false;
  static std::mutex mu;     // handles may be created from several host threads (SamplerPool lanes)
  std::lock_guard<std::mutex> lock(mu);
  if (done) return CFB_OK;
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, 3, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, 3, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, 3, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, 3, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, 3, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64, 4, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<64, 4>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<32, 4, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<32, 4>::TOTAL));
  if (const char* e = getenv("CFB_TC_CLUSTER")) g_tc_cluster = atoi(e);
  if (const char* e = getenv("CFB_TC_TMA_EPI")) g_tc_tma_epi = atoi(e);
  if (const char* e = getenv("CFB_TC_OCC3")) g_tc_occ3 = atoi(e);
  if (const char* e = getenv("CFB_TC_2CTA")) g_tc_pair = atoi(e);
  if (const char* e = getenv("CFB_TC_2CTA_MIN_ROWS")) g_tc_pair_min_rows = atoi(e);
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<128, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 2>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<64, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<64, 2>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<32, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<32, 2>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<128, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<64, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<64, 4>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<32, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<32, 4>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_grouped_tma_kernel<128, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 2>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_grouped_tma_kernel<128, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_grouped_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_ln_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  done = true;
  return CFB_OK;
}

bool gemm_tc_supported(int M, int N, int K, int lda, int ldw) {
  return M > 0 && K >= BK && K % BK == 0 && N % 32 == 0 && lda % 8 == 0 && ldw % 8 == 0;
}

static int check_epilogue(Epilogue& ep) {
  if (ep.replicate < 1) ep.replicate = 1;
  if (ep.bias_period < 1) ep.bias_period = 1;
  CFB_CHECK(!(ep.accumulate && ep.out_bf16), "gemm: accumulate needs float output");
  CFB_CHECK(ep.ldo % 8 == 0 && (ep.rep_stride % 8 == 0), "gemm_tc: ldo %% 8 must be 0");
  return CFB_OK;
}

int gemm_tc(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep_in,
            cudaStream_t st, int w_rows) {
  if (debug_skip(16)) return CFB_OK;
  CFB_CHECK(gemm_tc_supported(M, N, K, lda, ldw), "gemm_tc: unsupported shape %dx%dx%d", M, N, K);
  CFB_CHECK(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "gemm_tc: operands must be 16-byte aligned");
  Epilogue ep = ep_in;
  CFB_TRY(check_epilogue(ep));
  CFB_CHECK((uintptr_t)ep.out % 16 == 0, "gemm_tc: output must be 16-byte aligned");
  CFB_CHECK(ep.bias == nullptr || ((uintptr_t)ep.bias % 16 == 0), "gemm_tc: bias must be 16-byte aligned");
  if (ep.ln_out) {
    CFB_CHECK(N == 512 && ep.ldo == 512 && !ep.out_bf16 && ep.replicate == 1 && ep.ln_g && ep.ln_b,
              "gemm_tc: fused LayerNorm needs a float [M,512] output");
    if (ep.ln_counters == nullptr) return launch_ln(A, lda, W, ldw, M, N, K, ep, st);   // cluster / DSMEM variant
  }
  if (w_rows <= 0 || w_rows > N) w_rows = N;
  if (ep.ln_out && ep.ln_tail) return launch<128, 3, 1, 1>(A, lda, W, ldw, w_rows, M, N, K, ep, st);   // tail needs plain CTAs
  if (g_tc_pair && N % 256 == 0 && M >= g_tc_pair_min_rows && ep.ln_out == nullptr && tma_epilogue_ok(ep, 256))
    return launch_pair(A, lda, W, ldw, w_rows, M, N, K, ep, st);
  if (N % 128 == 0) {
    // cluster shape: g_tc_cluster = 10*CN_max + CM_max (env CFB_TC_CLUSTER); CN must divide the number of n-tiles.
    // Multicast pays when several tiles share an operand; a single m-tile (tiny M) gains nothing from CM.
    const int n_tiles = N / 128, m_tiles = ceil_div(M, BM);
    const int cn_max = g_tc_cluster / 10, cm_max = g_tc_cluster % 10;
    const int cn = (cn_max >= 4 && n_tiles % 4 == 0) ? 4 : ((cn_max >= 2 && n_tiles % 2 == 0) ? 2 : 1);
    const int cm = (cm_max >= 2 && m_tiles >= 2 && w_rows == N) ? 2 : 1;
    if (cn == 4 && cm == 2) return launch<128, 3, 4, 2>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
    if (cn == 2 && cm == 2) return launch<128, 3, 2, 2>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
    if (cn == 4) return launch<128, 3, 4, 1>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
    if (cn == 2) return launch<128, 3, 2, 1>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
    return launch<128, 3, 1, 1>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
  }
  if (N % 64 == 0) return launch<64, 4, 1, 1>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
  return launch<32, 4, 1, 1>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
}

int gemm_tc_grouped(const TcGroup* groups, int n_groups, int a_rows_total, int lda, int ldw, int N, int K,
                    const Epilogue& ep_in, cudaStream_t st) {
  if (debug_skip(32)) return CFB_OK;
  CFB_CHECK(n_groups > 0 && n_groups <= TC_MAX_GROUPS, "gemm_tc_grouped: %d groups (max %d)", n_groups, TC_MAX_GROUPS);
  CFB_CHECK(N % 128 == 0 && K % BK == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_tc_grouped: unsupported shape N=%d K=%d", N, K);
  Epilogue ep = ep_in;
  CFB_TRY(check_epilogue(ep));
  for (int z = 0; z < n_groups; ++z)
    CFB_CHECK(((uintptr_t)groups[z].A % 16 == 0) && ((uintptr_t)groups[z].W % 16 == 0) && ((uintptr_t)groups[z].out % 16 == 0) &&
              ((uintptr_t)groups[z].bias % 16 == 0), "gemm_tc_grouped: group %d operands must be 16-byte aligned", z);
  return launch_grouped<128, 3>(groups, n_groups, a_rows_total, lda, ldw, N, K, ep, st);
}

}  // namespace cfb
