// Attention kernels.
//  * mha_kernel: the core of torch.nn.MultiheadAttention on already-projected q/k/v (denoiser
//    self-attention 16x16 / 4 heads x 128; VAE self-attention 128x128 and cross-attention 128x8,
//    2 heads x 64) with an optional per-sample key length (key_padding_mask of lengths_to_mask).
//  * cross_kernel: the denoiser's five single-head cross-attentions in folded form: the query side
//    already carries W_k (qx = A_x norm2(tgt) + a_x), keys AND values are the affine-free
//    normalised memory rows, so one block per (batch entry, stream) computes
//    softmax(qx . mem_hat^T) . mem_hat over head_dim = d_model = 512.
#include <type_traits>
#include "common.cuh"
#include "kernels.cuh"
#include <cstdlib>
#include <mutex>

namespace cfb {

namespace {

template <typename T>
__global__ void __launch_bounds__(128) mha_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k,
                                                  const T* __restrict__ v, int ldk, T* __restrict__ out, int ldo,
                                                  int Lq, int Lk, int hd, const int* __restrict__ kv_len,
                                                  int q_per_block) {
  pdl_sync();
  extern __shared__ float sm[];
  const int b = blockIdx.x, h = blockIdx.y, q0 = blockIdx.z * q_per_block;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Ks = sm;                       // [Lk][hd+1]
  float* Vs = Ks + (size_t)Lk * (hd + 1);   // [Lk][hd]
  float* Qs = Vs + (size_t)Lk * hd;     // [4][hd]
  float* Ps = Qs + 4 * hd;              // [4][Lk]
  const int valid = kv_len ? min(kv_len[b], Lk) : Lk;
  for (int i = threadIdx.x; i < Lk * hd; i += blockDim.x) {
    const int j = i / hd, d = i % hd;
    const size_t g = (size_t)(b * Lk + j) * ldk + h * hd + d;
    Ks[j * (hd + 1) + d] = to_f32<T>(k[g]);
    Vs[j * hd + d] = to_f32<T>(v[g]);
  }
  __syncthreads();
  const float scale = sqrtf(1.0f / (float)hd);   // torch: q * math.sqrt(1.0 / head_dim)
  float* qs = Qs + warp * hd;
  float* ps = Ps + warp * Lk;
  const int q_end = min(q0 + q_per_block, Lq);
  for (int qi = q0 + warp; qi < q_end; qi += 4) {
    const T* qrow = q + (size_t)(b * Lq + qi) * ldq + h * hd;
    for (int d = lane; d < hd; d += 32) qs[d] = to_f32<T>(qrow[d]) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) {
      float s;
      if (j < valid) {
        s = 0.f;
        const float* kr = Ks + j * (hd + 1);
        for (int d = 0; d < hd; ++d) s = fmaf(qs[d], kr[d], s);
      } else {
        s = -INFINITY;
      }
      ps[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float e = expf(ps[j] - mx);
      ps[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    T* orow = out + (size_t)(b * Lq + qi) * ldo + h * hd;
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < valid; ++j) acc = fmaf(ps[j], Vs[j * hd + d], acc);
      orow[d] = from_f32<T>(acc * inv);
    }
    __syncwarp();
  }
}

constexpr int CROSS_D = 512;
constexpr int CROSS_MAXQ = 16;

template <typename T>
__device__ __forceinline__ void load_row16(const T* __restrict__ row, int lane, float v[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (sizeof(T) == 4) {
      float4 t = *reinterpret_cast<const float4*>(row + i * 128 + lane * 4);
      v[i * 4] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
    } else if constexpr (std::is_same<T, __half>::value) {
      uint2 t = *reinterpret_cast<const uint2*>(row + i * 128 + lane * 4);
      const float2 a = __half22float2(*reinterpret_cast<__half2*>(&t.x));
      const float2 b = __half22float2(*reinterpret_cast<__half2*>(&t.y));
      v[i * 4] = a.x; v[i * 4 + 1] = a.y; v[i * 4 + 2] = b.x; v[i * 4 + 3] = b.y;
    } else {
      uint2 t = *reinterpret_cast<const uint2*>(row + i * 128 + lane * 4);
      __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
      __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
      v[i * 4] = __low2float(a); v[i * 4 + 1] = __high2float(a);
      v[i * 4 + 2] = __low2float(b); v[i * 4 + 3] = __high2float(b);
    }
  }
}

// grid (n_batch, n_streams, 16 / QPB); 256 threads.  qx/u rows are [n_batch * 16, 5 * 512].
// A block owns QPB = 4 of the 16 queries of one (batch entry, stream) pair: the pair's work is CUDA-core fp32 math
// (about 5 MFLOP for the 161 audio keys) and splitting it four ways is what keeps all SMs busy when only one
// branch of the guidance batch is conditional on a stream.
// Shared memory: Qs [QPB][512] float queries, St [M][QPB] float scores/probabilities (key-major).
constexpr int QPB = 4;

template <typename T>
__global__ void __launch_bounds__(256) cross_kernel(const T* __restrict__ qx, const T* __restrict__ mem_hat,
                                                    T* __restrict__ u, CrossArgs a, int n_tokens) {
  pdl_sync();
  extern __shared__ float sm[];
  const int bs = blockIdx.x + a.bs_offset, x = blockIdx.y, q0 = blockIdx.z * QPB;
  const int M = a.len[x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = CFB_N_STREAMS * CROSS_D;
  float* Qs = sm;                   // [QPB][512]
  float* St = Qs + QPB * CROSS_D;   // [M][QPB]
  const int slot = a.slot[x] ? a.slot[x][bs] : bs;
  if (a.skip_slot0 && slot == 0) return;   // block-uniform: covered by the shared-slot GEMM path
  const T* mem = mem_hat + ((size_t)a.row_base[x] + (size_t)slot * M) * CROSS_D;
  const uint8_t* msk = a.mask[x] ? a.mask[x] + (size_t)slot * M : nullptr;

  for (int i = threadIdx.x * 4; i < QPB * CROSS_D; i += blockDim.x * 4) {   // stage the queries (zero past n_tokens)
    const int qi = q0 + i / CROSS_D, c = i % CROSS_D;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (qi < n_tokens) {
      const T* src = qx + (size_t)(bs * n_tokens + qi) * ld + x * CROSS_D + c;
      if constexpr (sizeof(T) == 4) {
        v = *reinterpret_cast<const float4*>(src);
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(src);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
        v = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
      }
    }
    *reinterpret_cast<float4*>(Qs + i) = v;
  }
  __syncthreads();
  // ---- scores: warp w walks keys w, w+8, ...; lanes split the 512 columns; the QPB partial sums are independent
  // chains, reduced across lanes by a transpose-reduce.
  float kv_next[16];
  if (warp < M) load_row16<T>(mem + (size_t)warp * CROSS_D, lane, kv_next);
  for (int j = warp; j < M; j += 8) {
    float kv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) kv[i] = kv_next[i];
    if (j + 8 < M) load_row16<T>(mem + (size_t)(j + 8) * CROSS_D, lane, kv_next);   // next key row in flight
    float s[QPB];
#pragma unroll
    for (int qi = 0; qi < QPB; ++qi) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 qv = *reinterpret_cast<const float4*>(Qs + qi * CROSS_D + i * 128 + lane * 4);
        acc = fmaf(qv.x, kv[i * 4], acc); acc = fmaf(qv.y, kv[i * 4 + 1], acc);
        acc = fmaf(qv.z, kv[i * 4 + 2], acc); acc = fmaf(qv.w, kv[i * 4 + 3], acc);
      }
      s[qi] = acc;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {     // lanes with bit 4 set keep queries 2,3
      const bool up = lane & 16;
      const float give = up ? s[i] : s[i + 2];
      const float got = __shfl_xor_sync(0xffffffffu, give, 16);
      s[i] = (up ? s[i + 2] : s[i]) + got;
    }
    {                                 // lanes with bit 3 set keep the odd query of their pair
      const bool up = lane & 8;
      const float give = up ? s[0] : s[1];
      const float got = __shfl_xor_sync(0xffffffffu, give, 8);
      s[0] = (up ? s[1] : s[0]) + got;
    }
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 4);
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 2);
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
    const int qsel = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    if ((lane & 7) == 0) St[j * QPB + qsel] = (msk && msk[j]) ? -INFINITY : s[0];
  }
  __syncthreads();
  // ---- softmax per query (warps 0..QPB-1)
  if (warp < QPB && q0 + warp < n_tokens) {
    const int ql = warp, qi = q0 + warp;
    float mx = -INFINITY;
    for (int j = lane; j < M; j += 32) mx = fmaxf(mx, St[j * QPB + ql]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < M; j += 32) {
      const float e = expf(St[j * QPB + ql] - mx);
      St[j * QPB + ql] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float* arow = nullptr;
    if (a.att[x] && bs >= a.att_first_batch) {
      const long long step_off = a.step_ptr ? (long long)(*a.step_ptr) * a.att_step_stride[x] : 0;
      arow = a.att[x] + step_off + (long long)(bs - a.att_first_batch) * a.att_batch_stride[x] + (long long)qi * M;
    }
    for (int j = lane; j < M; j += 32) {
      const float p = St[j * QPB + ql] * inv;
      St[j * QPB + ql] = p;
      if (arow) arow[j] = p;
    }
  }
  __syncthreads();
  // ---- u = P . mem_hat: thread owns columns {2t, 2t+1}; 8 key rows are fetched before any is consumed
  const int c = threadIdx.x * 2;
  float acc[QPB][2];
#pragma unroll
  for (int qi = 0; qi < QPB; ++qi) acc[qi][0] = acc[qi][1] = 0.f;
  constexpr int PF = 8;
  for (int j0 = 0; j0 < M; j0 += PF) {
    float v0[PF], v1[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int j = min(j0 + k, M - 1);
      if constexpr (sizeof(T) == 4) {
        const float2 t = *reinterpret_cast<const float2*>(mem + (size_t)j * CROSS_D + c);
        v0[k] = t.x; v1[k] = t.y;
      } else {
        const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(mem + (size_t)j * CROSS_D + c);
        v0[k] = __low2float(t); v1[k] = __high2float(t);
      }
    }
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int j = j0 + k;
      if (j < M) {
        const float4 p = *reinterpret_cast<const float4*>(St + j * QPB);   // warp-wide broadcast of the 4 probabilities
        acc[0][0] = fmaf(p.x, v0[k], acc[0][0]); acc[0][1] = fmaf(p.x, v1[k], acc[0][1]);
        acc[1][0] = fmaf(p.y, v0[k], acc[1][0]); acc[1][1] = fmaf(p.y, v1[k], acc[1][1]);
        acc[2][0] = fmaf(p.z, v0[k], acc[2][0]); acc[2][1] = fmaf(p.z, v1[k], acc[2][1]);
        acc[3][0] = fmaf(p.w, v0[k], acc[3][0]); acc[3][1] = fmaf(p.w, v1[k], acc[3][1]);
      }
    }
  }
#pragma unroll   // fully unrolled so acc[][] stays in registers
  for (int ql = 0; ql < QPB; ++ql) {
    const int qi = q0 + ql;
    if (qi >= n_tokens) break;
    T* o = u + (size_t)(bs * n_tokens + qi) * ld + x * CROSS_D + c;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float2*>(o) = make_float2(acc[ql][0], acc[ql][1]);
    } else {
      *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(acc[ql][0], acc[ql][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core version of the per-pair attention (bf16): one block per (batch entry, stream), 8 warps.
//   scores  S[16, M] = Q[16,512] . xhat^T : mma.sync.m16n8k16, A from shared memory (ldmatrix), B fragments
//           straight from the row-major key rows in global memory (a key row IS a column-major B column);
//           warp w owns key tiles w, w+8, ...
//   softmax in fp32 in shared memory -> P (bf16, zero padded to a multiple of 16 keys)
//   values  U[16,512] = P[16,M] . xhat[M,512] : key tiles of 16 rows are staged with cp.async (double buffered)
//           and read with ldmatrix.trans; warp w owns output columns [64w, 64w+64).
// 16-query tiles are exactly one MMA row block, so nothing is wasted on padding.
constexpr int XP = CROSS_D + 8;   // bf16 row pitch in shared memory: rows shift by 16 B -> conflict-free ldmatrix
constexpr int CM_NBUF = 4;        // value-tile ring of the tensor-core cross-attention: three 16-key tiles in flight
constexpr int QP = CROSS_D + 32;  // query rows shift by 64 B -> conflict-free 128-bit loads per quarter-warp (rows g, g+1 x 4 lanes)

__device__ __forceinline__ void ldsm_x4(uint32_t r[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t r[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16_16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 16-bit operands of either format: bf16, or fp16 when F16 (same fragment layouts)
template <bool F16>
__device__ __forceinline__ void mma_16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  if constexpr (F16) mma_f16_16816(c, a, b0, b1);
  else mma_bf16_16816(c, a, b0, b1);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// F16: qx and mem_hat hold fp16, the probabilities are rounded to fp16, both products run as f16 MMAs
template <bool F16>
__global__ void __launch_bounds__(256) cross_mma_kernel(const bf16* __restrict__ qx, const bf16* __restrict__ mem_hat,
                                                        bf16* __restrict__ u, CrossArgs a, int n_tokens, int Sp, int Pp) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int bs = blockIdx.x + a.bs_offset, x = blockIdx.y;
  const int M = a.len[x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ld = CFB_N_STREAMS * CROSS_D;
  const int slot = a.slot[x] ? a.slot[x][bs] : bs;
  if (a.skip_slot0 && slot == 0) return;   // block-uniform
  bf16* Qs = reinterpret_cast<bf16*>(smraw);            // [16][QP]
  bf16* Xs = Qs + 16 * QP;                              // [CM_NBUF][16][XP]
  float* Ss = reinterpret_cast<float*>(Xs + CM_NBUF * 16 * XP);   // [16][Sp]
  bf16* Ps = reinterpret_cast<bf16*>(Ss + 16 * Sp);     // [16][Pp]
  const bf16* mem = mem_hat + ((size_t)a.row_base[x] + (size_t)slot * M) * CROSS_D;
  const uint8_t* msk = a.mask[x] ? a.mask[x] + (size_t)slot * M : nullptr;
  const uint32_t qs_addr = (uint32_t)__cvta_generic_to_shared(Qs);
  const uint32_t xs_addr = (uint32_t)__cvta_generic_to_shared(Xs);
  const uint32_t ps_addr = (uint32_t)__cvta_generic_to_shared(Ps);

  // stage Q (zero rows past n_tokens): 16 rows x 64 chunks of 16 B
  for (int i = threadIdx.x; i < 16 * 64; i += 256) {
    const int r = i >> 6, ch = i & 63;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < n_tokens) v = *reinterpret_cast<const uint4*>(qx + (size_t)(bs * n_tokens + r) * ld + x * CROSS_D + ch * 8);
    *reinterpret_cast<uint4*>(Qs + r * QP + ch * 8) = v;
  }
  // first value tile in flight while the scores are computed
  const int nkt = (M + 15) >> 4;
  auto stage_tile = [&](int kt, int buf) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int r = i >> 6, ch = i & 63;
      const int j = min(kt * 16 + r, M - 1);
      cp_async16(xs_addr + (uint32_t)(((buf * 16 + r) * XP + ch * 8) * 2), mem + (size_t)j * CROSS_D + ch * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // the first value tiles stream in while the scores are computed (a group is committed per tile, empty past the end)
#pragma unroll
  for (int i = 0; i < CM_NBUF - 1; ++i) {
    if (i < nkt) stage_tile(i, i);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __syncthreads();

  // ---- scores.  The reduction index of an mma is free to be permuted as long as both operands agree, so lane t
  // takes the eight CONTIGUOUS columns 32 j + 8 t .. + 7 of its key row (one 128-bit load; two k-steps' worth of B
  // fragments) and the matching eight columns of query rows g and g + 8 from shared memory.  All 16 key loads of an
  // 8-key tile are in flight at once: one L2 round trip per tile instead of four.
  const int nnt = (M + 7) >> 3;
  for (int nt = warp; nt < nnt; nt += 8) {
    const int key0 = nt * 8;
    const bf16* kp = mem + (size_t)min(key0 + g, M - 1) * CROSS_D + 8 * t;
    uint4 kb[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) kb[j] = *reinterpret_cast<const uint4*>(kp + 32 * j);
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint4 qa = *reinterpret_cast<const uint4*>(Qs + g * QP + 32 * j + 8 * t);
      const uint4 qb = *reinterpret_cast<const uint4*>(Qs + (g + 8) * QP + 32 * j + 8 * t);
      const uint32_t a0[4] = {qa.x, qb.x, qa.y, qb.y}, a1[4] = {qa.z, qb.z, qa.w, qb.w};
      mma_16816<F16>(c, a0, kb[j].x, kb[j].y);
      mma_16816<F16>(c, a1, kb[j].z, kb[j].w);
    }
    const int j0 = key0 + 2 * t, j1 = j0 + 1;
    const bool m0 = j0 >= M || (msk && msk[j0]), m1 = j1 >= M || (msk && msk[min(j1, M - 1)]);
    Ss[g * Sp + j0] = m0 ? -INFINITY : c[0];
    Ss[g * Sp + j1] = m1 ? -INFINITY : c[1];
    Ss[(g + 8) * Sp + j0] = m0 ? -INFINITY : c[2];
    Ss[(g + 8) * Sp + j1] = m1 ? -INFINITY : c[3];
  }
  __syncthreads();
  // ---- softmax: warp w handles query rows w and w + 8; P is written as bf16, zero beyond M
  for (int qi = warp; qi < 16; qi += 8) {
    float* srow = Ss + qi * Sp;
    bf16* prow = Ps + qi * Pp;
    if (qi >= n_tokens) {
      for (int j = lane; j < nkt * 16; j += 32) prow[j] = __float2bfloat16_rn(0.f);   // zero in either format
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j < M; j += 32) mx = fmaxf(mx, srow[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < M; j += 32) {
      const float e = expf(srow[j] - mx);
      srow[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float* arow = nullptr;
    if (a.att[x] && bs >= a.att_first_batch) {
      const long long step_off = a.step_ptr ? (long long)(*a.step_ptr) * a.att_step_stride[x] : 0;
      arow = a.att[x] + step_off + (long long)(bs - a.att_first_batch) * a.att_batch_stride[x] + (long long)qi * M;
    }
    for (int j = lane; j < nkt * 16; j += 32) {
      float p = 0.f;
      if (j < M) {
        p = srow[j] * inv;
        if (arow) arow[j] = p;
      }
      if constexpr (F16) reinterpret_cast<__half*>(prow)[j] = __float2half_rn(p);
      else prow[j] = __float2bfloat16_rn(p);
    }
  }
  // ---- values
  float acc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt % CM_NBUF;
    asm volatile("cp.async.wait_group %0;" ::"n"(CM_NBUF - 2) : "memory");   // tiles 0..kt+NBUF-2 committed: tile kt is in
    __syncthreads();                         // tile kt landed for everyone; tile kt-1's buffer is free; P is visible
    if (kt + CM_NBUF - 1 < nkt) stage_tile(kt + CM_NBUF - 1, (kt + CM_NBUF - 1) % CM_NBUF);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    uint32_t af[4];
    ldsm_x4(af, ps_addr + (uint32_t)(((lane & 15) * Pp + kt * 16 + (lane >> 4) * 8) * 2));
    const int brow = (lane & 7) + ((lane >> 3) & 1) * 8;       // key row inside the tile
    const int bcol = warp * 64 + (lane >> 4) * 8;              // first of the two 8-column blocks of this x4
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      uint32_t bf[4];
      ldsm_x4_trans(bf, xs_addr + (uint32_t)(((buf * 16 + brow) * XP + bcol + nn * 16) * 2));
      mma_16816<F16>(acc[2 * nn], af, bf[0], bf[1]);
      mma_16816<F16>(acc[2 * nn + 1], af, bf[2], bf[3]);
    }
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int col = x * CROSS_D + warp * 64 + n * 8 + 2 * t;
    if (g < n_tokens)
      *reinterpret_cast<uint32_t*>(u + (size_t)(bs * n_tokens + g) * ld + col) = pack16(acc[n][0], acc[n][1], a.out_f16);
    if (g + 8 < n_tokens)
      *reinterpret_cast<uint32_t*>(u + (size_t)(bs * n_tokens + g + 8) * ld + col) = pack16(acc[n][2], acc[n][3], a.out_f16);
  }
}


// Tensor-core multi-head attention core for bf16 (mma.sync.m16n8k16, fp32 accumulation): one warp owns 16 query rows
// of one (sample, head); K, V (zero-padded to NKT 16-key tiles) and the block's Q rows are staged once in shared
// memory with cp.async.  Scores stay in registers: the accumulator layout of S = Q K^T is exactly the A-fragment
// layout of the second product, so softmax(S) is normalised, rounded to bf16 and fed to P V without leaving the
// warp.  Used for the denoiser self-attention (16 x 16, head_dim 128) and the VAE attentions (head_dim 64; 128 x 128
// self, 128 x 8 cross, 18 x 18 encoder).
// F16: q, k, v hold fp16 (written by an fp16-output GEMM epilogue), the probabilities are rounded to fp16 and both
// products run as f16 MMAs; out_f16 selects the output format (bf16 for the denoiser's out_proj, fp16 in the fp16 VAE).
template <int HD, int NKT, bool F16 = false>
__global__ void __launch_bounds__(128) mha_mma_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ k,
                                                      const bf16* __restrict__ v, int ldk, bf16* __restrict__ out, int ldo,
                                                      int Lq, int Lk, const int* __restrict__ kv_len, float scale,
                                                      int out_f16) {
  pdl_sync();
  constexpr int P = HD + 8;                    // bf16 row pitch: rows shift by 16 B -> conflict-free ldmatrix
  constexpr int CH = HD / 8;                   // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smraw[];
  const int b = blockIdx.x, h = blockIdx.y, q0 = blockIdx.z * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nq = blockDim.x >> 1;              // query rows staged by this block (16 per warp)
  bf16* Ks = reinterpret_cast<bf16*>(smraw);   // [NKT*16][P]
  bf16* Vs = Ks + NKT * 16 * P;
  bf16* Qs = Vs + NKT * 16 * P;                // [nq][P]
  const uint32_t ks_addr = (uint32_t)__cvta_generic_to_shared(Ks);
  const uint32_t vs_addr = (uint32_t)__cvta_generic_to_shared(Vs);
  const uint32_t qs_addr = (uint32_t)__cvta_generic_to_shared(Qs);
  const int valid = kv_len ? min(kv_len[b], Lk) : Lk;
  for (int i = threadIdx.x; i < NKT * 16 * CH; i += blockDim.x) {
    const int r = i / CH, ch = i % CH;
    if (r < Lk) {
      cp_async16(ks_addr + (uint32_t)((r * P + ch * 8) * 2), k + (size_t)(b * Lk + r) * ldk + h * HD + ch * 8);
      cp_async16(vs_addr + (uint32_t)((r * P + ch * 8) * 2), v + (size_t)(b * Lk + r) * ldk + h * HD + ch * 8);
    } else {
      *reinterpret_cast<uint4*>(Ks + r * P + ch * 8) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Vs + r * P + ch * 8) = make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = threadIdx.x; i < nq * CH; i += blockDim.x) {
    const int r = i / CH, ch = i % CH;
    if (q0 + r < Lq) cp_async16(qs_addr + (uint32_t)((r * P + ch * 8) * 2), q + (size_t)(b * Lq + q0 + r) * ldq + h * HD + ch * 8);
    else *reinterpret_cast<uint4*>(Qs + r * P + ch * 8) = make_uint4(0, 0, 0, 0);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int row0 = q0 + warp * 16;
  if (row0 >= Lq) return;
  // ---- S = Q K^T
  float s[NKT * 2][4];
#pragma unroll
  for (int n = 0; n < NKT * 2; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t af[4];
    ldsm_x4(af, qs_addr + (uint32_t)(((warp * 16 + (lane & 15)) * P + kk * 16 + (lane >> 4) * 8) * 2));
#pragma unroll
    for (int j = 0; j < NKT; ++j) {
      uint32_t bf[4];
      ldsm_x4(bf, ks_addr + (uint32_t)(((j * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * P + kk * 16 + ((lane >> 3) & 1) * 8) * 2));
      mma_16816<F16>(s[2 * j], af, bf[0], bf[1]);
      mma_16816<F16>(s[2 * j + 1], af, bf[2], bf[3]);
    }
  }
  // ---- softmax over the valid keys (rows g and g + 8 of this warp's tile; a row lives in one quad)
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NKT * 2; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool dead = n * 8 + 2 * t + e >= valid;
      s[n][e] = dead ? -INFINITY : s[n][e] * scale;
      s[n][2 + e] = dead ? -INFINITY : s[n][2 + e] * scale;
      mx0 = fmaxf(mx0, s[n][e]);
      mx1 = fmaxf(mx1, s[n][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int n = 0; n < NKT * 2; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[n][e] = expf(s[n][e] - mx0); sum0 += s[n][e];
      s[n][2 + e] = expf(s[n][2 + e] - mx1); sum1 += s[n][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  // ---- O = P V
  float o[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int j = 0; j < NKT; ++j) {
    uint32_t af[4];
    {
      af[0] = pack16(s[2 * j][0] * inv0, s[2 * j][1] * inv0, F16);
      af[1] = pack16(s[2 * j][2] * inv1, s[2 * j][3] * inv1, F16);
      af[2] = pack16(s[2 * j + 1][0] * inv0, s[2 * j + 1][1] * inv0, F16);
      af[3] = pack16(s[2 * j + 1][2] * inv1, s[2 * j + 1][3] * inv1, F16);
    }
    const int brow = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int nn = 0; nn < HD / 16; ++nn) {
      uint32_t bf[4];
      ldsm_x4_trans(bf, vs_addr + (uint32_t)((brow * P + nn * 16 + (lane >> 4) * 8) * 2));
      mma_16816<F16>(o[2 * nn], af, bf[0], bf[1]);
      mma_16816<F16>(o[2 * nn + 1], af, bf[2], bf[3]);
    }
  }
  const int r_lo = row0 + g, r_hi = row0 + g + 8;
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    const int col = h * HD + n * 8 + 2 * t;
    if (r_lo < Lq) *reinterpret_cast<uint32_t*>(out + (size_t)(b * Lq + r_lo) * ldo + col) = pack16(o[n][0], o[n][1], out_f16);
    if (r_hi < Lq) *reinterpret_cast<uint32_t*>(out + (size_t)(b * Lq + r_hi) * ldo + col) = pack16(o[n][2], o[n][3], out_f16);
  }
}

template <int HD, int NKT, bool F16 = false>
int launch_mha_mma(const bf16* q, int ldq, const bf16* k, const bf16* v, int ldk, bf16* out, int ldo, int n, int Lq, int Lk,
                   int n_heads, const int* kv_len, cudaStream_t st, int out_f16 = 0) {
  const int warps = Lq >= 64 ? 4 : ceil_div(Lq, 16);
  const size_t smem = (size_t)(2 * NKT * 16 + warps * 16) * (HD + 8) * 2;
  dim3 grid(n, n_heads, ceil_div(Lq, 64));
  launch_k(mha_mma_kernel<HD, NKT, F16>, grid, warps * 32, smem, st, q, ldq, k, v, ldk, out, ldo, Lq, Lk, kv_len,
           sqrtf(1.0f / (float)HD), out_f16);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

constexpr int ATT_MAX_SMEM = 160 * 1024;
int g_mha_simt = 0;   // env CFB_MHA_SIMT=1: CUDA-core attention kernels in bf16 mode too

// Denoiser self-attention (cross_attention.py:570): 16 tokens, head_dim 128.  One single-warp block owns one
// (sample, head) -- 25 KB of shared memory, so the blocks slot in beside resident GEMM CTAs instead of waiting for a
// whole SM; lane = (query i, parity p).  q/k/v of the head are staged as float in shared memory
// with a 132-float row pitch so every 128-bit access below is conflict-free per quarter-warp: lane
// (i,p) scores keys j = 2*jj+p (adjacent rows -> banks +4) and accumulates output chunks 2*c+p.
constexpr int SA_L = 16, SA_HD = 128, SA_PITCH = 132;

template <typename T>
__global__ void __launch_bounds__(32) self_attn16_kernel(const T* __restrict__ qkv, int ld, int E,
                                                         T* __restrict__ out, int ldo, int n_heads) {
  pdl_sync();
  extern __shared__ float sm[];
  const int lane = threadIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.x;
  float* Q = sm;
  float* K = Q + SA_L * SA_PITCH;
  float* V = K + SA_L * SA_PITCH;
  // 48 rows of 128 elements, 4 per lane (coalesced); 12 rows are in flight per round trip
  constexpr int RBATCH = 12;
#pragma unroll 1
  for (int r0 = 0; r0 < 3 * SA_L; r0 += RBATCH) {
    float4 v[RBATCH];
#pragma unroll
    for (int k = 0; k < RBATCH; ++k) {
      const int r = r0 + k, which = r / SA_L, row = r % SA_L;
      const T* src = qkv + (size_t)(b * SA_L + row) * ld + which * E + h * SA_HD + lane * 4;
      if constexpr (sizeof(T) == 4) {
        v[k] = *reinterpret_cast<const float4*>(src);
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(src);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
        const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
        v[k] = make_float4(__low2float(a), __high2float(a), __low2float(c), __high2float(c));
      }
    }
#pragma unroll
    for (int k = 0; k < RBATCH; ++k) {
      const int r = r0 + k, which = r / SA_L, row = r % SA_L;
      if (which == 0) {                          // torch scales q by sqrt(1/head_dim) before q k^T
        const float s = sqrtf(1.0f / (float)SA_HD);
        v[k].x *= s; v[k].y *= s; v[k].z *= s; v[k].w *= s;
      }
      *reinterpret_cast<float4*>(Q + which * SA_L * SA_PITCH + row * SA_PITCH + lane * 4) = v[k];
    }
  }
  __syncwarp();
  const int i = lane >> 1, p = lane & 1;
  float s[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) s[jj] = 0.f;
#pragma unroll 4
  for (int d = 0; d < SA_HD; d += 4) {
    const float4 qv = *reinterpret_cast<const float4*>(Q + i * SA_PITCH + d);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float4 kv = *reinterpret_cast<const float4*>(K + (2 * jj + p) * SA_PITCH + d);
      s[jj] = fmaf(qv.x, kv.x, s[jj]); s[jj] = fmaf(qv.y, kv.y, s[jj]);
      s[jj] = fmaf(qv.z, kv.z, s[jj]); s[jj] = fmaf(qv.w, kv.w, s[jj]);
    }
  }
  float mx = s[0];
#pragma unroll
  for (int jj = 1; jj < 8; ++jj) mx = fmaxf(mx, s[jj]);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  float sum = 0.f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) { s[jj] = expf(s[jj] - mx); sum += s[jj]; }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = 1.0f / sum;
  float4 acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const float mine = s[jj] * inv;
    const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
    const float p_even = p ? other : mine;     // key 2*jj
    const float p_odd = p ? mine : other;      // key 2*jj + 1
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 v0 = *reinterpret_cast<const float4*>(V + (2 * jj) * SA_PITCH + (2 * c + p) * 4);
      const float4 v1 = *reinterpret_cast<const float4*>(V + (2 * jj + 1) * SA_PITCH + (2 * c + p) * 4);
      acc[c].x = fmaf(p_even, v0.x, acc[c].x); acc[c].y = fmaf(p_even, v0.y, acc[c].y);
      acc[c].z = fmaf(p_even, v0.z, acc[c].z); acc[c].w = fmaf(p_even, v0.w, acc[c].w);
      acc[c].x = fmaf(p_odd, v1.x, acc[c].x); acc[c].y = fmaf(p_odd, v1.y, acc[c].y);
      acc[c].z = fmaf(p_odd, v1.z, acc[c].z); acc[c].w = fmaf(p_odd, v1.w, acc[c].w);
    }
  }
  T* orow = out + (size_t)(b * SA_L + i) * ldo + h * SA_HD;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    T* o = orow + (2 * c + p) * 4;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(o) = acc[c];
    } else {
      const __nv_bfloat162 a = __floats2bfloat162_rn(acc[c].x, acc[c].y);
      const __nv_bfloat162 d2 = __floats2bfloat162_rn(acc[c].z, acc[c].w);
      uint2 pk; pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&d2);
      *reinterpret_cast<uint2*>(o) = pk;
    }
  }
}

constexpr int SA_SMEM = 3 * SA_L * SA_PITCH * (int)sizeof(float);   // 25,344 B per one-warp block: fits beside two GEMM CTAs

// One warp per query row; see SharedAttnArgs.
template <typename TP>
__global__ void __launch_bounds__(256) softmax_shared_kernel(const float* __restrict__ S, TP* __restrict__ P,
                                                             SharedAttnArgs a, int rows, int n_tokens) {
  pdl_sync();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int bs = r / n_tokens + a.bs_offset;
#pragma unroll
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    TP* prow = P + (size_t)r * a.ld_p + a.p_off[x];
    const int M = a.len[x], kp = a.kp[x];
    if (a.slot[x][bs] != 0) {
      for (int j = lane; j < kp; j += 32) prow[j] = from_f32<TP>(0.f);
      continue;
    }
    const float* srow = S + (size_t)r * a.ld_s + a.s_off[x];
    const uint8_t* msk = a.mask[x];
    float mx = -INFINITY;
    for (int j = lane; j < M; j += 32) {
      const float s = (msk && msk[j]) ? -INFINITY : srow[j];
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < M; j += 32) {
      const float s = (msk && msk[j]) ? -INFINITY : srow[j];
      sum += expf(s - mx);
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < kp; j += 32) {
      float p = 0.f;
      if (j < M) {
        const float s = (msk && msk[j]) ? -INFINITY : srow[j];
        p = expf(s - mx) * inv;
      }
      prow[j] = from_f32<TP>(p);
    }
  }
}


// Register-resident variant: one warp per (query row, stream).  The stream's score segment (<= 32 * SMX_MAXJ keys)
// is fetched in ONE round trip together with the slot id and the mask bytes, the softmax runs out of registers and
// every probability is written once; five times as many warps as rows hide what latency is left.  The strided
// kernel above walks each segment three times with dependent loads (max, sum, write): ~15 L2 round trips per row,
// 12-15 us per launch for about a microsecond of work.
constexpr int SMX_MAXJ = 8;
template <typename TP>
__global__ void __launch_bounds__(256) softmax_shared_reg_kernel(const float* __restrict__ S, TP* __restrict__ P,
                                                                 SharedAttnArgs a, int rows, int n_tokens) {
  pdl_sync();
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int r = w / CFB_N_STREAMS, x = w % CFB_N_STREAMS;
  if (r >= rows) return;
  const int bs = r / n_tokens + a.bs_offset;
  const int M = a.len[x], kp = a.kp[x];
  const float* srow = S + (size_t)r * a.ld_s + a.s_off[x];
  TP* prow = P + (size_t)r * a.ld_p + a.p_off[x];
  const uint8_t* msk = a.mask[x];
  const int slot = a.slot[x][bs];
  float s[SMX_MAXJ];
#pragma unroll
  for (int i = 0; i < SMX_MAXJ; ++i) {
    const int j = lane + 32 * i;
    s[i] = -INFINITY;
    if (j < M) {
      const float v = srow[j];
      const bool dead = msk && msk[j];
      s[i] = dead ? -INFINITY : v;
    }
  }
  float mx = s[0];
#pragma unroll
  for (int i = 1; i < SMX_MAXJ; ++i) mx = fmaxf(mx, s[i]);
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < SMX_MAXJ; ++i) {
    if (32 * i < M) { s[i] = expf(s[i] - mx); sum += s[i]; }   // warp-uniform test; masked / out-of-range -> exp(-inf) = 0
  }
  sum = warp_sum(sum);
  const float inv = slot == 0 ? 1.0f / sum : 0.f;   // conditional pairs are handled by cross_attention(): zeros here
#pragma unroll
  for (int i = 0; i < SMX_MAXJ; ++i) {
    const int j = lane + 32 * i;
    if (j < kp) prow[j] = from_f32<TP>(j < M ? s[i] * inv : 0.f);
  }
  for (int j = 32 * SMX_MAXJ + lane; j < kp; j += 32) prow[j] = from_f32<TP>(0.f);
}

// z0[l][s_off[x] + j] = a_{x,l} . xhat_{x,slot 0, j}: the key-dependent bias of the shared-slot scores.
struct Z0Args {
  const float* a_zx[CFB_N_STREAMS];   // [L, 512]
  int row_base[CFB_N_STREAMS], len[CFB_N_STREAMS], s_off[CFB_N_STREAMS];
  int tok_base[CFB_N_STREAMS + 1];    // prefix sum of len
};
template <typename T>
__global__ void __launch_bounds__(256) z0_kernel(const T* __restrict__ mem_hat, float* __restrict__ z0, Z0Args a,
                                                 int n_layers, int n_tot) {
  pdl_sync();
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= a.tok_base[CFB_N_STREAMS]) return;
  int x = 0;
#pragma unroll
  for (int i = 1; i < CFB_N_STREAMS; ++i) x += (t >= a.tok_base[i]);
  const int j = t - a.tok_base[x];
  float kv[16];
  load_row16<T>(mem_hat + ((size_t)a.row_base[x] + j) * CROSS_D, lane, kv);
  for (int l = 0; l < n_layers; ++l) {
    const float* av = a.a_zx[x] + (size_t)l * CROSS_D;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 q = *reinterpret_cast<const float4*>(av + i * 128 + lane * 4);
      s = fmaf(q.x, kv[i * 4], s); s = fmaf(q.y, kv[i * 4 + 1], s);
      s = fmaf(q.z, kv[i * 4 + 2], s); s = fmaf(q.w, kv[i * 4 + 3], s);
    }
    s = warp_sum(s);
    if (lane == 0) z0[(size_t)l * n_tot + a.s_off[x] + j] = s;
  }
}

}  // namespace

template <typename T>
int shared_key_bias(const T* mem_hat, float* z0, const float* const a_zx[CFB_N_STREAMS],
                    const int row_base[CFB_N_STREAMS], const int len[CFB_N_STREAMS], const int s_off[CFB_N_STREAMS],
                    int n_layers, int n_tot, cudaStream_t st, int f16) {
  Z0Args a;
  int tok = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    a.a_zx[x] = a_zx[x]; a.row_base[x] = row_base[x]; a.len[x] = len[x]; a.s_off[x] = s_off[x];
    a.tok_base[x] = tok; tok += len[x];
  }
  a.tok_base[CFB_N_STREAMS] = tok;
  if (f16 && sizeof(T) == 2)
    launch_k(z0_kernel<__half>, ceil_div(tok, 8), 256, 0, st, reinterpret_cast<const __half*>(mem_hat), z0, a, n_layers, n_tot);
  else launch_k(z0_kernel<T>, ceil_div(tok, 8), 256, 0, st, mem_hat, z0, a, n_layers, n_tot);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int shared_key_bias<bf16>(const bf16*, float*, const float* const*, const int*, const int*, const int*, int, int, cudaStream_t, int);
template int shared_key_bias<float>(const float*, float*, const float* const*, const int*, const int*, const int*, int, int, cudaStream_t, int);

template <typename TP>
int softmax_shared(const float* S, TP* P, const SharedAttnArgs& a, int n_batch, int n_tokens, cudaStream_t st) {
  const int rows = n_batch * n_tokens;
  if (rows <= 0 || debug_skip(4)) return CFB_OK;
  static const bool strided = getenv("CFB_SOFTMAX_STRIDED") && atoi(getenv("CFB_SOFTMAX_STRIDED"));
  bool fits = true;
  for (int x = 0; x < CFB_N_STREAMS; ++x) fits = fits && a.len[x] <= 32 * SMX_MAXJ;
  if constexpr (sizeof(TP) == 2) {
    if (a.p_f16) {     // same layout, fp16 payload
      __half* Ph = reinterpret_cast<__half*>(P);
      if (fits && !strided) launch_k(softmax_shared_reg_kernel<__half>, ceil_div(rows * CFB_N_STREAMS, 8), 256, 0, st, S, Ph, a, rows, n_tokens);
      else launch_k(softmax_shared_kernel<__half>, ceil_div(rows, 8), 256, 0, st, S, Ph, a, rows, n_tokens);
      CFB_LAUNCH_CHECK();
      return CFB_OK;
    }
  }
  if (fits && !strided) launch_k(softmax_shared_reg_kernel<TP>, ceil_div(rows * CFB_N_STREAMS, 8), 256, 0, st, S, P, a, rows, n_tokens);
  else launch_k(softmax_shared_kernel<TP>, ceil_div(rows, 8), 256, 0, st, S, P, a, rows, n_tokens);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int softmax_shared<bf16>(const float*, bf16*, const SharedAttnArgs&, int, int, cudaStream_t);
template int softmax_shared<float>(const float*, float*, const SharedAttnArgs&, int, int, cudaStream_t);

namespace {
}  // namespace

// Opt in to large dynamic shared memory once PER DEVICE (the attribute is per device), outside any stream capture.
int init_attention_kernels() {
  static unsigned long long done_mask = 0;
  static std::mutex mu;     // handles may be created from several host threads (SamplerPool lanes)
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  CFB_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && ((done_mask >> dev) & 1ull)) return CFB_OK;
  if (const char* e = getenv("CFB_MHA_SIMT")) g_mha_simt = atoi(e);
  CFB_CUDA(cudaFuncSetAttribute(mha_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(mha_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(mha_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(cross_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(cross_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(cross_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(cross_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(self_attn16_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SA_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(self_attn16_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SA_SMEM));
  if (dev < 64) done_mask |= 1ull << dev;
  return CFB_OK;
}

// Self-attention on fp16 q / k / v (the denoiser's 16 x 16, head_dim 128 case on the mma.sync kernel); bf16 or fp16 output.
bool mha_f16_supported(int Lk, int head_dim) {
  return head_dim == 128 && Lk <= 16 && g_gemm_backend != CFB_GEMM_SIMT && !g_mha_simt;
}
int mha_f16(const bf16* q, int ldq, const bf16* k, const bf16* v, int ldk, bf16* out, int ldo, int n, int Lq, int Lk,
            int n_heads, int head_dim, cudaStream_t st, int out_f16) {
  if (n <= 0 || debug_skip(2)) return CFB_OK;
  const bool aligned = ldq % 8 == 0 && ldk % 8 == 0 && ldo % 2 == 0 && ((uintptr_t)q % 16 == 0) &&
                       ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 4 == 0);
  CFB_CHECK(aligned && mha_f16_supported(Lk, head_dim), "mha_f16: unsupported (Lk=%d head_dim=%d)", Lk, head_dim);
  return launch_mha_mma<128, 1, true>(q, ldq, k, v, ldk, out, ldo, n, Lq, Lk, n_heads, nullptr, st, out_f16);
}

template <typename T>
int mha(const T* q, int ldq, const T* k, const T* v, int ldk, T* out, int ldo, int n, int Lq, int Lk, int n_heads,
        int head_dim, const int* kv_len, cudaStream_t st) {
  if (n <= 0 || debug_skip(2)) return CFB_OK;
  CFB_CHECK(Lq > 0 && Lk > 0 && head_dim > 0 && n_heads > 0, "mha: bad shape");
  constexpr bool HF = std::is_same<T, __half>::value;   // fp16 VAE: f16 MMAs, fp16 output
  if constexpr (sizeof(T) == 2) {   // 16-bit: tensor-core kernel (CFB_GEMM_SIMT keeps the CUDA-core engines for cross-checks)
    const bool aligned = ldq % 8 == 0 && ldk % 8 == 0 && ldo % 2 == 0 && ((uintptr_t)q % 16 == 0) &&
                         ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 4 == 0);
    if (aligned && g_gemm_backend != CFB_GEMM_SIMT && !g_mha_simt) {
      const bf16 *q_ = (const bf16*)q, *k_ = (const bf16*)k, *v_ = (const bf16*)v;
      bf16* o_ = (bf16*)out;
      if (head_dim == 128 && Lk <= 16) return launch_mha_mma<128, 1, HF>(q_, ldq, k_, v_, ldk, o_, ldo, n, Lq, Lk, n_heads, kv_len, st, HF);
      if (head_dim == 64 && Lk <= 16) return launch_mha_mma<64, 1, HF>(q_, ldq, k_, v_, ldk, o_, ldo, n, Lq, Lk, n_heads, kv_len, st, HF);
      if (head_dim == 64 && Lk <= 32) return launch_mha_mma<64, 2, HF>(q_, ldq, k_, v_, ldk, o_, ldo, n, Lq, Lk, n_heads, kv_len, st, HF);
      if (head_dim == 64 && Lk <= 128) return launch_mha_mma<64, 8, HF>(q_, ldq, k_, v_, ldk, o_, ldo, n, Lq, Lk, n_heads, kv_len, st, HF);
    }
  }
  // packed self-attention of the denoiser: q, k, v are column blocks of one [rows, 3E] matrix
  if constexpr (!HF) {   // (its 16-bit loads / stores are bf16; the fp16 element type takes the generic kernel below)
    if (Lq == SA_L && Lk == SA_L && head_dim == SA_HD && kv_len == nullptr && ldq == ldk &&
        k == q + n_heads * head_dim && v == q + 2 * n_heads * head_dim && ldq % 4 == 0 && ldo % 4 == 0) {
      dim3 grid(n, n_heads);
      launch_k(self_attn16_kernel<T>, grid, 32, SA_SMEM, st, q, ldq, n_heads * head_dim, out, ldo, n_heads);
      CFB_LAUNCH_CHECK();
      return CFB_OK;
    }
  }
  const size_t smem = ((size_t)Lk * (head_dim + 1) + (size_t)Lk * head_dim + 4 * head_dim + 4 * Lk) * sizeof(float);
  CFB_CHECK(smem <= (size_t)ATT_MAX_SMEM, "mha: Lk=%d head_dim=%d needs %zu B of shared memory", Lk, head_dim, smem);
  const int q_per_block = Lq >= 64 ? 32 : Lq;
  dim3 grid(n, n_heads, ceil_div(Lq, q_per_block));
  launch_k(mha_kernel<T>, grid, 128, smem, st, q, ldq, k, v, ldk, out, ldo, Lq, Lk, head_dim, kv_len, q_per_block);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int mha<float>(const float*, int, const float*, const float*, int, float*, int, int, int, int, int, int, const int*, cudaStream_t);
template int mha<bf16>(const bf16*, int, const bf16*, const bf16*, int, bf16*, int, int, int, int, int, int, const int*, cudaStream_t);
template int mha<__half>(const __half*, int, const __half*, const __half*, int, __half*, int, int, int, int, int, int, const int*, cudaStream_t);

template <typename T>
int cross_attention(const T* qx, const T* mem_hat, T* u, const CrossArgs& a, int n_batch, int n_tokens, int d,
                    cudaStream_t st) {
  if (n_batch <= 0 || debug_skip(8)) return CFB_OK;
  CFB_CHECK(d == CROSS_D && n_tokens <= CROSS_MAXQ, "cross_attention: d=%d n_tokens=%d unsupported", d, n_tokens);
  int maxM = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    CFB_CHECK(a.len[x] > 0, "cross_attention: stream %d has no memory tokens", x);
    if (a.len[x] > maxM) maxM = a.len[x];
  }
  if constexpr (sizeof(T) == 2) {   // bf16: tensor-core kernels (in-place u == qx is fine: Q is staged before u is written)
    if (!a.out_f16 && !a.in_f16 && cross_tc_supported(a, n_tokens, d))   // tcgen05 / TMEM / TMA (cross_tc.cu); else mma.sync below
      return cross_attention_tc(qx, (a.bs_offset + n_batch) * n_tokens, mem_hat, u, a, n_batch, st);
    const int Sp = ((maxM + 7) & ~7) + 8, Pp = ((maxM + 15) & ~15) + 8;
    const size_t smem_mma = (size_t)(16 * QP + CM_NBUF * 16 * XP) * 2 + (size_t)16 * Sp * 4 + (size_t)16 * Pp * 2;
    // CFB_GEMM_SIMT selects the CUDA-core engines everywhere (tests cross-check the two implementations)
    if (smem_mma <= (size_t)ATT_MAX_SMEM && g_gemm_backend != CFB_GEMM_SIMT) {
      dim3 grid(n_batch, CFB_N_STREAMS);
      if (a.in_f16) launch_k(cross_mma_kernel<true>, grid, 256, smem_mma, st, qx, mem_hat, u, a, n_tokens, Sp, Pp);
      else launch_k(cross_mma_kernel<false>, grid, 256, smem_mma, st, qx, mem_hat, u, a, n_tokens, Sp, Pp);
      CFB_LAUNCH_CHECK();
      return CFB_OK;
    }
  }
  CFB_CHECK(!a.out_f16 && !a.in_f16, "cross_attention: fp16 operands exist for the mma.sync kernel only (%d memory tokens; "
            "cfb_set_bf16_activation_f16(0) selects bf16 outputs)", maxM);
  const size_t smem = ((size_t)QPB * CROSS_D + (size_t)QPB * maxM) * sizeof(float);
  CFB_CHECK(smem <= (size_t)ATT_MAX_SMEM, "cross_attention: %d memory tokens exceed the shared-memory budget", maxM);
  dim3 grid(n_batch, CFB_N_STREAMS, ceil_div(n_tokens, QPB));
  launch_k(cross_kernel<T>, grid, 256, smem, st, qx, mem_hat, u, a, n_tokens);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int cross_attention<float>(const float*, const float*, float*, const CrossArgs&, int, int, int, cudaStream_t);
template int cross_attention<bf16>(const bf16*, const bf16*, bf16*, const CrossArgs&, int, int, int, cudaStream_t);

}  // namespace cfb
