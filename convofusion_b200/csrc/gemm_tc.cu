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
#include "tc_ptx.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <cstring>
#include <cstdlib>

namespace cfb {

namespace {

using namespace tc;

constexpr int BM = 128;      // UMMA M (cta_group::1)
constexpr int BK = 64;       // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 192;

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

template <int BN, int STAGES, bool A2 = false>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = (A2 ? 2 : 1) * A_BYTES + B_BYTES;   // A2: the hi and the lo tile of the A operand
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int STATS_OFF = BAR_OFF + 128;                             // BN bias floats
  static constexpr int TOTAL = BAR_OFF + 128 + 1024 + 1024;                   // barriers, bias, alignment slack
};

// One 128 x BN output tile: rows [m0, min(m0+128, m_end)), columns [n0, n0+BN).
// TMA_ONLY selects the TMA store / L2 reduce-add epilogue at compile time (fewer registers: the staged ld/st epilogue
// is not even in the binary); otherwise ep.tma_out picks at run time.
// Variants that were measured slower on this workload and removed in round 2 (DESIGN.md 5.1 keeps the numbers): cluster
// TMA multicast of the operands, a 2-stage ring with three CTAs per SM, and three ways of running the following
// LayerNorm inside this kernel (last-arriving CTA, 4-CTA cluster over DSMEM, a spin-waiting tail).
// A2: the A operand carries TWO bf16 terms per value (hi + lo: activations at 16 mantissa bits),
// stored as [hi(64) | lo(64)] per 64 columns (row stride 2 K): a stage holds both tiles and one W tile, and every K step
// issues two accumulating MMAs (hi . w + lo . w) -- twice the MMAs on an idle tensor pipe, +50 % operand bytes.
template <int BN, int STAGES, bool TMA_ONLY = false, bool A2 = false>
__device__ __forceinline__ void gemm_tile(const CUtensorMap* tmA_p, const CUtensorMap* tmB_p, const CUtensorMap* tmO_p, const int m0,
                                          const int M /* first row NOT to store */, const int n0, const int N,
                                          const int K, const Epilogue& ep) {
  using S = Smem<BN, STAGES, A2>;
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
      mbar_init(bar_empty + s * 8, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  pdl_sync();   // everything above touched only shared/tensor memory; operands of the previous kernel are read below
  if (threadIdx.x == 0) trace(1);

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_empty + s * 8, ph ^ 1);
        const uint32_t sa = base + s * S::STAGE_BYTES, sb = sa + (A2 ? 2 : 1) * S::A_BYTES;
        mbar_expect_tx(bar_full + s * 8, S::STAGE_BYTES);
        tma_load_2d(sa, &tmA, (A2 ? 2 * kb : kb) * BK, m0, bar_full + s * 8);
        if constexpr (A2) tma_load_2d(sa + S::A_BYTES, &tmA, (2 * kb + 1) * BK, m0, bar_full + s * 8);
        tma_load_2d(sb, &tmB, kb * BK, n0, bar_full + s * 8);
        if (kb == 0) trace(2);
      }
      trace(3);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // bits 7 / 10 cleared = A and B hold fp16 (Epilogue::ab_f16; A = f16 with B = bf16 traps as an illegal instruction)
      const uint32_t idesc = make_idesc(BM, BN) & ~(ep.ab_f16 ? ((1u << 7) | (1u << 10)) : 0u);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        tc_fence_after();
        if (kb == 0) trace(4);
        const uint32_t sa = base + s * S::STAGE_BYTES, sb = sa + (A2 ? 2 : 1) * S::A_BYTES;
        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advancing K inside the 128-byte swizzle row: +32 bytes = +2 in the (addr >> 4) field
          umma_bf16(tmem_acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          if constexpr (A2)
            umma_bf16(tmem_acc, make_smem_desc(sa + S::A_BYTES) + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
        }
        umma_commit(bar_empty + s * 8);   // frees the smem slot once these MMAs retire
      }
      umma_commit(bar_acc);
      trace(5);
    }
  } else {
    // ---- epilogue.  tcgen05.ld hands each thread one accumulator ROW, which is the wrong shape for global memory
    // (a warp would touch 32 different cache lines per instruction).  The pipeline stages are idle once bar_acc has
    // fired, so each warp transposes its 32 x BN slab through that shared memory and then walks its rows with the
    // lanes spread across the columns: every global access of the bias / residual / output is row-contiguous.
    const int q = warp & 3;
    if (TMA_ONLY || ep.tma_out) {
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
              for (int e = 0; e < 4; ++e) pk[e] = pack16(v[8 * j + 2 * e], v[8 * j + 2 * e + 1], ep.out_f16);
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
        if (lane == 0) tma_wait_read0();   // the boxes must stay intact until the TMA unit has read them
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
              uint2 pk;
              pk.x = pack16(v[i][0], v[i][1], ep.out_f16); pk.y = pack16(v[i][2], v[i][3], ep.out_f16);
              *reinterpret_cast<uint2*>(o) = pk;
            } else if constexpr (CPL == 2) {
              *reinterpret_cast<uint32_t*>(o) = pack16(v[i][0], v[i][1], ep.out_f16);
            } else {
              *reinterpret_cast<uint16_t*>(o) = (uint16_t)(pack16(v[i][0], 0.f, ep.out_f16) & 0xffffu);
            }
          }
        }
      }
    }
    }   // staged (non-TMA) epilogue
  }
  if (warp == 2 && lane == 0) trace(8);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace(9);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, BN);
    if (lane == 0) trace(10);
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmO, int M, int N,
                                                           int K, Epilogue ep) {
  gemm_tile<BN, STAGES>(&tmA, &tmB, &tmO, blockIdx.y * BM, M, blockIdx.x * BN, N, K, ep);
}

// TMA-epilogue-only instantiation: no staged ld/st epilogue in the binary (116 registers, 2 CTAs per SM): the phases
// of co-resident CTAs (prologue, operand latency, main loop, epilogue) overlap on one SM.
template <int BN, int STAGES, bool A2 = false>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tc_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmO, int M, int N,
                                                               int K, Epilogue ep) {
  gemm_tile<BN, STAGES, true, A2>(&tmA, &tmB, &tmO, blockIdx.y * BM, M, blockIdx.x * BN, N, K, ep);
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
// ONE pair CTA per SM (4-stage ring = 130 KB of shared memory): tcgen05.alloc.cta_group::2 is a collective of the pair that
// holds each SM's allocation permit until both CTAs have joined.  With two pair CTAs co-resident per SM, pair A's CTA on
// SM0 and pair B's CTA on SM1 can each hold their SM's permit while their peers wait for the other one -- the round-1
// hang with two batches in flight (pairs of different kernels interleaved on one TPC).  Single-CTA kernels next to a pair
// CTA are harmless: cta_group::1 allocations never wait for another SM.
template <int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA,
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
  cluster_sync_all();                                // both CTAs of the pair are resident before either asks for tensor memory
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
      const uint32_t idesc = make_idesc(256, BN2) & ~(ep.ab_f16 ? ((1u << 7) | (1u << 10)) : 0u);
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
            for (int e = 0; e < 4; ++e) pk[e] = pack16(v[8 * j + 2 * e], v[8 * j + 2 * e + 1], ep.out_f16);
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

template <int BN, int STAGES, bool A2 = false>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tc_grouped_tma_kernel(const __grid_constant__ GroupedArgs g, int N,
                                                                          int K, Epilogue ep) {
  const int z = blockIdx.z;
  const int m0 = g.row_start[z] + blockIdx.y * BM;
  const int m_end = g.row_start[z] + g.rows[z];
  if (m0 >= m_end) return;
  ep.bias = g.bias[z];
  ep.out = g.out[z];
  gemm_tile<BN, STAGES, true, A2>(&g.tmA[z], &g.tmB[z], &g.tmO[z], m0, m_end, blockIdx.x * BN, N, K, ep);
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

// Descriptors are pure functions of (pointer, shape), so they are cached; the cache is bounded (callers hand in fresh
// torch allocations on the eager paths): when it fills up it is simply dropped and rebuilt on demand.
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
  if (cache.size() >= 4096) cache.clear();
  cache.emplace(key, tm);
  *out = tm;
  return CFB_OK;
}

int g_tc_tma_epi = 1;   // env CFB_TC_TMA_EPI=0 keeps the shared-memory-staged ld/st epilogue

// The TMA epilogue takes one output copy and (for bf16) tiles at least one 64-column box wide.
bool tma_epilogue_ok(const Epilogue& ep, int BN_) {
  return g_tc_tma_epi && ep.replicate == 1 && (BN_ >= 64 || !ep.out_bf16);
}

template <int BN, int STAGES>
int launch(const bf16* A, int lda, const bf16* W, int ldw, int w_rows, int M, int N, int K, const Epilogue& ep_in,
           cudaStream_t st) {
  using S = Smem<BN, STAGES>;
  CUtensorMap ta, tb, to;
  Epilogue ep = ep_in;
  CFB_TRY(get_map(A, M, K, lda, BM, &ta));
  CFB_TRY(get_map(W, w_rows, K, ldw, BN, &tb));     // rows past w_rows read as zeros (TMA out-of-bounds fill)
  ep.tma_out = tma_epilogue_ok(ep, BN);
  if (ep.tma_out) CFB_TRY(get_map(ep.out, M, N, ep.ldo, 32, &to, ep.out_bf16 ? MAP_OUT_BF16 : MAP_OUT_F32));
  else to = ta;
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  if (ep.tma_out) launch_k(gemm_tc_tma_kernel<BN, STAGES>, grid, NTHREADS, S::TOTAL, st, ta, tb, to, M, N, K, ep);
  else launch_k(gemm_tc_kernel<BN, STAGES>, grid, NTHREADS, S::TOTAL, st, ta, tb, to, M, N, K, ep);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

// A operand with two bf16 terms per value ([hi | lo] per 64 columns, row stride lda >= 2 K): TMA epilogue only.
template <int BN, int STAGES>
int launch_a2(const bf16* A, int lda, const bf16* W, int ldw, int w_rows, int M, int N, int K, const Epilogue& ep_in,
              cudaStream_t st) {
  using S = Smem<BN, STAGES, true>;
  CUtensorMap ta, tb, to;
  Epilogue ep = ep_in;
  CFB_TRY(get_map(A, M, 2 * K, lda, BM, &ta));
  CFB_TRY(get_map(W, w_rows, K, ldw, BN, &tb));
  ep.tma_out = 1;
  CFB_TRY(get_map(ep.out, M, N, ep.ldo, 32, &to, ep.out_bf16 ? MAP_OUT_BF16 : MAP_OUT_F32));
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  launch_k(gemm_tc_tma_kernel<BN, STAGES, true>, grid, NTHREADS, S::TOTAL, st, ta, tb, to, M, N, K, ep);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int g_tc_pair = 0;            // env CFB_TC_2CTA: 1 = cta_group::2 256x256 tiles for N % 256 == 0
int g_tc_pair_min_rows = 0;   // env CFB_TC_2CTA_MIN_ROWS: only for GEMMs with at least this many rows

int launch_pair(const bf16* A, int lda, const bf16* W, int ldw, int w_rows, int M, int N, int K, const Epilogue& ep_in,
                cudaStream_t st) {
  using S = Smem<128, 4>;
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
  CFB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<4>, ta, tb, to, M, N, K, ep));
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

template <int BN, int STAGES, bool A2 = false>
int launch_grouped(const TcGroup* groups, int n_groups, int a_rows_total, int lda, int ldw, int N, int K,
                   const Epilogue& ep_in, cudaStream_t st) {
  using S = Smem<BN, STAGES, A2>;
  GroupedArgs g;
  memset(&g, 0, sizeof(g));
  Epilogue ep = ep_in;
  ep.tma_out = tma_epilogue_ok(ep, BN);
  if (A2) CFB_CHECK(ep.tma_out && lda >= 2 * K, "gemm_tc_grouped: two-term A operand needs the TMA epilogue and lda >= 2 K");
  int max_rows = 0;
  for (int z = 0; z < n_groups; ++z) {
    CFB_TRY(get_map(groups[z].A, a_rows_total, (A2 ? 2 : 1) * K, lda, BM, &g.tmA[z]));
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
  if constexpr (A2) {
    launch_k(gemm_tc_grouped_tma_kernel<BN, STAGES, true>, grid, NTHREADS, S::TOTAL, st, g, N, K, ep);
  } else {
    if (ep.tma_out)
      launch_k(gemm_tc_grouped_tma_kernel<BN, STAGES>, grid, NTHREADS, S::TOTAL, st, g, N, K, ep);
    else
      launch_k(gemm_tc_grouped_kernel<BN, STAGES>, grid, NTHREADS, S::TOTAL, st, g, N, K, ep);
  }
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

}  // namespace

// tensor maps for the other tcgen05 kernels (rowblock.cu): kind 0 = bf16 operand box 64 x box_rows, 1 = fp32 box
// 32 x box_rows, 2 = bf16 output box 64 x box_rows
int tc_get_map(const void* p, int rows, int cols, int ld, int box_rows, CUtensorMap* out, int kind) {
  return get_map(p, rows, cols, ld, box_rows, out, kind);
}

// the two-term A operand (Epilogue::a_terms == 2) is implemented by the TMA-epilogue kernels only
bool gemm_tc_two_term_ok() { return g_tc_tma_epi != 0; }

int tc_trace_read(unsigned long long out[16]) {
  CFB_CUDA(cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(unsigned long long) * 16));
  return CFB_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: opt in once per device (a process may hold
// handles on several GPUs), outside any stream capture.
int init_gemm_tc_kernels() {
  static std::mutex mu;     // handles may be created from several host threads (SamplerPool lanes)
  static unsigned long long done_mask = 0;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  CFB_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && ((done_mask >> dev) & 1ull)) return CFB_OK;
  if (const char* e = getenv("CFB_TC_TMA_EPI")) g_tc_tma_epi = atoi(e);
  if (const char* e = getenv("CFB_TC_2CTA")) g_tc_pair = atoi(e);
  if (const char* e = getenv("CFB_TC_2CTA_MIN_ROWS")) g_tc_pair_min_rows = atoi(e);
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<64, 4>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<32, 4>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 4>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<64, 4>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<32, 4>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_grouped_tma_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_grouped_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 3>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<128, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 2, true>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_tma_kernel<64, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<64, 2, true>::TOTAL));
  CFB_CUDA(cudaFuncSetAttribute(gemm_tc_grouped_tma_kernel<128, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, 2, true>::TOTAL));
  if (dev < 64) done_mask |= 1ull << dev;
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
  if (debug_skip(128) && ep_in.accumulate) return CFB_OK;      // diagnosis: drop only the residual-update GEMMs
  if (debug_skip(256) && !ep_in.accumulate) return CFB_OK;     // ... only the others
  CFB_CHECK(gemm_tc_supported(M, N, K, lda, ldw), "gemm_tc: unsupported shape %dx%dx%d", M, N, K);
  CFB_CHECK(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "gemm_tc: operands must be 16-byte aligned");
  Epilogue ep = ep_in;
  if (debug_skip(64)) ep.accumulate = 0;                       // diagnosis: plain TMA store instead of the L2 reduce-add
  CFB_TRY(check_epilogue(ep));
  CFB_CHECK((uintptr_t)ep.out % 16 == 0, "gemm_tc: output must be 16-byte aligned");
  CFB_CHECK(ep.bias == nullptr || ((uintptr_t)ep.bias % 16 == 0), "gemm_tc: bias must be 16-byte aligned");
  if (w_rows <= 0 || w_rows > N) w_rows = N;
  if (ep.a_terms == 2) {
    CFB_CHECK(lda >= 2 * K && N % 64 == 0 && tma_epilogue_ok(ep, N % 128 == 0 ? 128 : 64),
              "gemm_tc: two-term A operand needs lda >= 2 K, N %% 64 == 0 and the TMA epilogue (%dx%dx%d)", M, N, K);
    if (N % 128 == 0) return launch_a2<128, 2>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
    return launch_a2<64, 2>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
  }
  if (g_tc_pair && N % 256 == 0 && M >= g_tc_pair_min_rows && tma_epilogue_ok(ep, 256))
    return launch_pair(A, lda, W, ldw, w_rows, M, N, K, ep, st);
  if (N % 128 == 0) return launch<128, 3>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
  if (N % 64 == 0) return launch<64, 4>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
  return launch<32, 4>(A, lda, W, ldw, w_rows, M, N, K, ep, st);
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
  if (ep.a_terms == 2) return launch_grouped<128, 2, true>(groups, n_groups, a_rows_total, lda, ldw, N, K, ep, st);
  return launch_grouped<128, 3>(groups, n_groups, a_rows_total, lda, ldw, N, K, ep, st);
}

}  // namespace cfb
