// Per-pair cross-attention on the tcgen05 tensor cores (sm_100a): one CTA per conditional (batch entry, stream) pair.
//
// The denoiser's folded cross-attention has keys == values == xhat [len, 512] and only 16 queries
// per batch entry, below the M >= 64 of tcgen05.mma.  Both products are therefore issued TRANSPOSED, with the 16 queries
// as the N dimension (N = 16 is a legal UMMA shape, 8 cycles per K step) and the long dimension as M:
//     S^T [keys, 16]  = xhat   [keys (M = 128 per tile), 512 (K)] . Q^T      A = xhat tile,    B = Q [16, 512]  (K-major)
//     O^T [512, 16]   = xhat^T [dims (M = 4 x 128), keys (K)]     . P^T      A = xhat^T tile,  B = P [16, keys] (K-major)
// so every operand is K-major under the 128-byte swizzle -- the second product reads a transposed copy of the memory
// (mem_hat_t, written once per step next to mem_hat) instead of an MN-major descriptor.  TMA (3-D tensor maps
// [slot, row, col]: rows past the slot's length are zero-filled, not fetched) feeds a 4-stage ring of 16 KB tiles; the
// scores land in tensor memory with one KEY per lane, the softmax over the keys runs across lanes (warp shuffles + one
// shared-memory exchange between the four softmax warps), the probabilities go back to shared memory as the B operand
// of the second product (and to the attention-map output when it is requested), and the 512 x 16 result leaves tensor
// memory with one output DIMENSION per lane, i.e. coalesced bf16 rows of `u`.
//
// Alternative to cross_mma_kernel (mma.sync.m16n8k16 + cp.async, attention.cu) for bf16 handles (cross_attention.py:593-626),
// selected with CFB_CROSS_TC=1 / cfb_set_cross_tc(1).  NOT the default: measured on the B200 inside the concurrent step it
// takes 16.9 us per launch against 13.3 us for the mma.sync kernel (6.37 k vs 7.07 k motion-s/s for the whole pass; a
// 6-stage ring with one CTA per SM: 5.86 k).  A pair is 16 queries against <= 161 keys -- 5 MFLOP behind 330 KB of
// operand traffic that has to stream through the ring twice (scores, then values from the transposed copy) -- so the
// kernel is bound by dependent TMA round trips, and the M = 16 mma.sync tile with direct loads is the better fit.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 tensor-memory allocation + MMA issue, warps 2..5 softmax /
// epilogue (TMEM lane quarter = warp % 4).
#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <mutex>

namespace cfb {

int tc_get_map(const void* p, int rows, int cols, int ld, int box_rows, CUtensorMap* out, int kind);   // gemm_tc.cu

namespace {

using namespace tc;

constexpr int XT_D = 512, XT_Q = 16, XT_THREADS = 192, XT_STAGES = 4;
constexpr int XT_TILE = 128 * 128;            // 128 rows x 64 bf16 (16 KB), the A operand of both products
constexpr int XT_QBLK = XT_Q * 128;           // 16 rows x 64 bf16 (2 KB), the B operand blocks
constexpr int XT_MAX_KEYS = 256;              // 2 key tiles of 128 / 4 key blocks of 64
constexpr int OFF_Q = 0;                                  // 8 K blocks of Q
constexpr int OFF_P = OFF_Q + 8 * XT_QBLK;                // up to 4 key blocks of P
constexpr int OFF_RING = 32 * 1024;                       // 1024-aligned tiles
constexpr int OFF_BAR = OFF_RING + XT_STAGES * XT_TILE;
constexpr int OFF_RED = OFF_BAR + 256;
constexpr int XT_SMEM = OFF_RED + 2 * 4 * XT_Q * 4 + 1024;   // + alignment slack
static_assert(OFF_P + 4 * XT_QBLK <= OFF_RING, "cross_tc: operand blocks overlap the ring");

struct XtMaps {
  CUtensorMap q;                        // qx [rows, 5 * 512], box 16 x 64
  CUtensorMap k[CFB_N_STREAMS];         // mem_hat   of stream x as [n_slots, len, 512],  box 1 x 128 x 64
  CUtensorMap kt[CFB_N_STREAMS];        // mem_hat_t of stream x as [n_slots, 512, lenp], box 1 x 128 x 64
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Guarded wait: a protocol error traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void xt_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();            // ~2 s
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// grid (n_batch, 5); u / qx rows are [n_batch * 16, 5 * 512] (u may alias qx: Q is staged before u is written).
__global__ void __launch_bounds__(XT_THREADS, 2) cross_tc_kernel(const __grid_constant__ XtMaps maps, bf16* __restrict__ u,
                                                                 CrossArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bs = blockIdx.x + a.bs_offset, x = blockIdx.y;
  const int M = a.len[x];
  const int slot = a.slot[x] ? a.slot[x][bs] : bs;
  if (a.skip_slot0 && slot == 0) return;      // block-uniform, before any barrier / allocation: the shared-slot path
  const int nt = (M + 127) >> 7;              // key tiles of 128 (M dimension of the first product)
  const int nkb = (M + 63) >> 6;              // key blocks of 64 (K dimension of the second product)
  const uint32_t sQ = base + OFF_Q, sP = base + OFF_P, sRing = base + OFF_RING, bars = base + OFF_BAR;
  const uint32_t bar_q = bars, bar_full = bars + 8, bar_empty = bar_full + 8 * XT_STAGES,
                 bar_s = bar_empty + 8 * XT_STAGES, bar_p = bar_s + 8, bar_o = bar_p + 8, tmem_slot = bar_o + 8;
  float* red = reinterpret_cast<float*>(gen + OFF_RED);       // [2][4][16]

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.k[x]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.kt[x]) : "memory");
    mbar_init(bar_q, 1);
    for (int s = 0; s < XT_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + OFF_BAR + (tmem_slot - bars));
  const uint32_t tmem_s = tmem, tmem_o = tmem + 32;           // S^T: 2 tiles x 16 columns; O^T: 4 tiles x 16 columns
  pdl_sync();

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(bar_q, 8 * XT_QBLK);
      for (int kb = 0; kb < 8; ++kb) tma_load_2d(sQ + kb * XT_QBLK, &maps.q, x * XT_D + kb * 64, bs * XT_Q, bar_q);
      uint32_t it = 0;
      for (int t = 0; t < nt; ++t)
        for (int kb = 0; kb < 8; ++kb, ++it) {
          const uint32_t s = it % XT_STAGES;
          xt_wait(bar_empty + 8 * s, ((it / XT_STAGES) & 1u) ^ 1u);
          mbar_expect_tx(bar_full + 8 * s, XT_TILE);
          tma_load_3d(sRing + s * XT_TILE, &maps.k[x], kb * 64, t * 128, slot, bar_full + 8 * s);
        }
      for (int kb2 = 0; kb2 < nkb; ++kb2)
        for (int m = 0; m < 4; ++m, ++it) {
          const uint32_t s = it % XT_STAGES;
          xt_wait(bar_empty + 8 * s, ((it / XT_STAGES) & 1u) ^ 1u);
          mbar_expect_tx(bar_full + 8 * s, XT_TILE);
          tma_load_3d(sRing + s * XT_TILE, &maps.kt[x], kb2 * 64, m * 128, slot, bar_full + 8 * s);
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, XT_Q);
      xt_wait(bar_q, 0);
      tc_fence_after();
      uint32_t it = 0;
      for (int t = 0; t < nt; ++t)
        for (int kb = 0; kb < 8; ++kb, ++it) {
          const uint32_t s = it % XT_STAGES;
          xt_wait(bar_full + 8 * s, (it / XT_STAGES) & 1u);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(sRing + s * XT_TILE), bdesc = make_smem_desc(sQ + kb * XT_QBLK);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_s + (uint32_t)(t * XT_Q), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          umma_commit(bar_empty + 8 * s);
        }
      umma_commit(bar_s);
      xt_wait(bar_p, 0);                   // probabilities are in shared memory (and the scores have been read)
      tc_fence_after();
      for (int kb2 = 0; kb2 < nkb; ++kb2)
        for (int m = 0; m < 4; ++m, ++it) {
          const uint32_t s = it % XT_STAGES;
          xt_wait(bar_full + 8 * s, (it / XT_STAGES) & 1u);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(sRing + s * XT_TILE), bdesc = make_smem_desc(sP + kb2 * XT_QBLK);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_o + (uint32_t)(m * XT_Q), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb2 | k) != 0);
          umma_commit(bar_empty + 8 * s);
        }
      umma_commit(bar_o);
    }
  } else {
    // ------------------------------------------------------------------------------------------ softmax + epilogue
    const int q = warp & 3;                                    // TMEM lane quarter of this warp
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint8_t* msk = a.mask[x] ? a.mask[x] + (size_t)slot * M : nullptr;
    float* arow = nullptr;
    if (a.att[x] && bs >= a.att_first_batch) {
      const long long step_off = a.step_ptr ? (long long)(*a.step_ptr) * a.att_step_stride[x] : 0;
      arow = a.att[x] + step_off + (long long)(bs - a.att_first_batch) * a.att_batch_stride[x];
    }
    xt_wait(bar_s, 0);
    tc_fence_after();
    float v[2][XT_Q];
    bool valid[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = t * 128 + q * 32 + lane;
      valid[t] = t < nt && j < M && !(msk && msk[j]);
      if (t < nt) tmem_ld16(tmem_s + lane_off + (uint32_t)(t * XT_Q), v[t]);
#pragma unroll
      for (int c = 0; c < XT_Q; ++c) v[t][c] = valid[t] ? v[t][c] : -INFINITY;
    }
    // column maxima over all keys: lanes, then the four warps
    float mx[XT_Q];
#pragma unroll
    for (int c = 0; c < XT_Q; ++c) mx[c] = warp_max(fmaxf(v[0][c], v[1][c]));
    if (lane == 0) {                                           // (every lane holds all 16 maxima after the butterfly)
#pragma unroll
      for (int c = 0; c < XT_Q; ++c) red[q * XT_Q + c] = mx[c];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
    for (int c = 0; c < XT_Q; ++c)
      mx[c] = fmaxf(fmaxf(red[c], red[XT_Q + c]), fmaxf(red[2 * XT_Q + c], red[3 * XT_Q + c]));
    float sm_[XT_Q];
#pragma unroll
    for (int c = 0; c < XT_Q; ++c) {
#pragma unroll
      for (int t = 0; t < 2; ++t) v[t][c] = valid[t] ? expf(v[t][c] - mx[c]) : 0.f;
      sm_[c] = warp_sum(v[0][c] + v[1][c]);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < XT_Q; ++c) red[4 * XT_Q + q * XT_Q + c] = sm_[c];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
    for (int c = 0; c < XT_Q; ++c) {
      const float tot = (red[4 * XT_Q + c] + red[5 * XT_Q + c]) + (red[6 * XT_Q + c] + red[7 * XT_Q + c]);
      sm_[c] = 1.0f / tot;
    }
    // probabilities: B operand of the second product, [16 queries][64 keys] bf16 blocks under the 128-byte swizzle
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = t * 128 + q * 32 + lane;
      if (t < nt && j < nkb * 64) {
        uint8_t* blk = gen + OFF_P + (j >> 6) * XT_QBLK;
        const int jj = j & 63;
#pragma unroll
        for (int c = 0; c < XT_Q; ++c) {
          const float p = v[t][c] * sm_[c];
          *reinterpret_cast<bf16*>(blk + c * 128 + ((((jj >> 3) ^ (c & 7)) << 4) | ((jj & 7) << 1))) = __float2bfloat16_rn(p);
          if (arow && j < M) arow[(size_t)c * M + j] = p;
        }
      }
    }
    fence_async_smem();                    // generic-proxy writes of P -> tcgen05.mma reads
    tc_fence_before();
    mbar_arrive_cta(bar_p);
    // ---- O^T: one output dimension per lane, 16 queries per thread -> coalesced rows of u
    xt_wait(bar_o, 0);
    tc_fence_after();
    const int ld = CFB_N_STREAMS * XT_D;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float o[XT_Q];
      tmem_ld16(tmem_o + lane_off + (uint32_t)(m * XT_Q), o);
      bf16* dst = u + (size_t)(bs * XT_Q) * ld + x * XT_D + m * 128 + q * 32 + lane;
#pragma unroll
      for (int c = 0; c < XT_Q; ++c) dst[(size_t)c * ld] = __float2bfloat16_rn(o[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// mem_hat [n_slots, len, 512] -> mem_hat_t [n_slots, 512, lenp] (keys past len are written as zeros)
__global__ void __launch_bounds__(256) mem_transpose_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int len,
                                                            int lenp) {
  pdl_sync();
  __shared__ bf16 tile[32][34];
  const int s = blockIdx.z, r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i;
    tile[ty + 8 * i][tx] = r < len ? in[((size_t)s * len + r) * XT_D + c0 + tx] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int dim = c0 + ty + 8 * i, key = r0 + tx;
    if (key < lenp) out[((size_t)s * XT_D + dim) * lenp + key] = tile[tx][ty + 8 * i];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn xt_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 [n2, n1, n0] (n0 contiguous), box 64 x 128 x 1, 128-byte swizzle, zero fill outside
int map3(const void* p, int n0, int n1, int n2, CUtensorMap* out) {
  EncodeTiledFn enc = xt_encode();
  CFB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  cuuint64_t gdim[3] = {(cuuint64_t)n0, (cuuint64_t)n1, (cuuint64_t)n2};
  cuuint64_t gstr[2] = {(cuuint64_t)n0 * 2, (cuuint64_t)n0 * n1 * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed (%d) for ptr=%p dims=%d,%d,%d", (int)r, p, n0, n1, n2);
    return CFB_ERR_CUDA;
  }
  return CFB_OK;
}

}  // namespace

int g_cross_tc = 0;   // env CFB_CROSS_TC=1 / cfb_set_cross_tc: this kernel instead of the mma.sync per-pair attention

int init_cross_tc_kernels() {
  static std::mutex mu;
  static unsigned long long done_mask = 0;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  CFB_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && ((done_mask >> dev) & 1ull)) return CFB_OK;
  if (const char* e = getenv("CFB_CROSS_TC")) g_cross_tc = atoi(e);
  CFB_CUDA(cudaFuncSetAttribute(cross_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XT_SMEM));
  if (dev < 64) done_mask |= 1ull << dev;
  return CFB_OK;
}

bool cross_tc_supported(const CrossArgs& a, int n_tokens, int d) {
  if (!g_cross_tc || g_gemm_backend == CFB_GEMM_SIMT || n_tokens != XT_Q || d != XT_D || a.mem_hat_t == nullptr) return false;
  for (int x = 0; x < CFB_N_STREAMS; ++x)
    if (a.len[x] > XT_MAX_KEYS || a.n_slots[x] <= 0) return false;
  return true;
}

int mem_transpose(const bf16* mem_hat, bf16* mem_hat_t, const CrossArgs& a, cudaStream_t st) {
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    const int len = a.len[x], lenp = a.lenp[x];
    launch_k(mem_transpose_kernel, dim3(ceil_div(lenp, 32), XT_D / 32, a.n_slots[x]), dim3(256), 0, st,
             mem_hat + (size_t)a.row_base[x] * XT_D, mem_hat_t + a.t_off[x], len, lenp);
    CFB_LAUNCH_CHECK();
  }
  return CFB_OK;
}

int cross_attention_tc(const bf16* qx, int q_rows, const bf16* mem_hat, bf16* u, const CrossArgs& a, int n_batch,
                       cudaStream_t st) {
  XtMaps maps;
  CFB_TRY(tc_get_map(qx, q_rows, CFB_N_STREAMS * XT_D, CFB_N_STREAMS * XT_D, XT_Q, &maps.q, 0));
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    CFB_TRY(map3(mem_hat + (size_t)a.row_base[x] * XT_D, XT_D, a.len[x], a.n_slots[x], &maps.k[x]));
    CFB_TRY(map3(a.mem_hat_t + a.t_off[x], a.lenp[x], XT_D, a.n_slots[x], &maps.kt[x]));
  }
  launch_k(cross_tc_kernel, dim3(n_batch, CFB_N_STREAMS), dim3(XT_THREADS), XT_SMEM, st, maps, u, a);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

}  // namespace cfb
