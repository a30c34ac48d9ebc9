// One warp owns one row of D floats: 128-bit accesses, statistics by warp shuffle, two-pass variance (mean first,
// then the sum of squared deviations) like ATen's LayerNorm.  Shared by the row kernels (rowops.cu) and the LayerNorm
// tail of the tcgen05 GEMM (gemm_tc.cu), so both produce bit-identical rows.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace cfb {

constexpr float LN_EPS = 1e-5f;

template <int D>
struct RowVec {
  static constexpr int PER_LANE = D / 32;  // 16 (D=512) or 4 (D=128)
  static constexpr int NV = PER_LANE / 4;
  float v[PER_LANE];
  // lane owns columns {i*128 + lane*4 .. +3} for i in [0, NV): coalesced float4 accesses
  __device__ __forceinline__ void load(const float* __restrict__ row, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 t = *reinterpret_cast<const float4*>(row + i * 128 + lane * 4);
      v[i * 4] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
    }
  }
  // L2-only loads: rows another SM has just written within the same kernel (GEMM tail) must not come from L1
  __device__ __forceinline__ void load_cg(const float* row, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 t = __ldcg(reinterpret_cast<const float4*>(row + i * 128 + lane * 4));
      v[i * 4] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
    }
  }
  __device__ __forceinline__ void add(const float* __restrict__ row, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 t = *reinterpret_cast<const float4*>(row + i * 128 + lane * 4);
      v[i * 4] += t.x; v[i * 4 + 1] += t.y; v[i * 4 + 2] += t.z; v[i * 4 + 3] += t.w;
    }
  }
  __device__ __forceinline__ void normalize() {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) s += v[i];
    const float mu = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) { v[i] -= mu; q += v[i] * v[i]; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + LN_EPS);
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) v[i] *= rstd;
  }
  __device__ __forceinline__ void affine(const float* __restrict__ g, const float* __restrict__ b, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 gg = *reinterpret_cast<const float4*>(g + i * 128 + lane * 4);
      float4 bb = *reinterpret_cast<const float4*>(b + i * 128 + lane * 4);
      v[i * 4] = v[i * 4] * gg.x + bb.x; v[i * 4 + 1] = v[i * 4 + 1] * gg.y + bb.y;
      v[i * 4 + 2] = v[i * 4 + 2] * gg.z + bb.z; v[i * 4 + 3] = v[i * 4 + 3] * gg.w + bb.w;
    }
  }
  template <typename T>
  __device__ __forceinline__ void store(T* __restrict__ row, int lane) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(row + i * 128 + lane * 4) = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
      } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[i * 4], v[i * 4 + 1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[i * 4 + 2], v[i * 4 + 3]);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(row + i * 128 + lane * 4) = pk;
      }
    }
  }
  // fp16 instead of bf16 (11 instead of 8 significant bits; the values are LayerNorm outputs, far inside the fp16 range,
  // and are clamped to it anyway): the A operand of a tcgen05 GEMM whose instruction descriptor says A = f16 (gemm_tc.cu).
  __device__ __forceinline__ void store_f16(bf16* __restrict__ row, int lane) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint2 pk;
      pk.x = f16x2_sat(v[i * 4], v[i * 4 + 1]);
      pk.y = f16x2_sat(v[i * 4 + 2], v[i * 4 + 3]);
      *reinterpret_cast<uint2*>(row + i * 128 + lane * 4) = pk;
    }
  }
  // Two bf16 terms per value (hi = bf16(v), lo = bf16(v - hi): 16 mantissa bits), laid out [hi(64) | lo(64)] per 64
  // columns with a row stride of 2 D -- the A operand of the tcgen05 GEMM's two-term mode (gemm_tc.cu, A2).
  // F16: both terms are fp16 (hi = fp16(v), lo = fp16(v - hi): 22 significant bits) for an all-f16 GEMM.
  template <bool F16 = false>
  __device__ __forceinline__ void store_split(bf16* __restrict__ row, int lane) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      bf16* dst = row + (c >> 6) * 128 + (c & 63);
      uint2 hi, lo;
      if constexpr (F16) {
        hi.x = f16x2_sat(v[i * 4], v[i * 4 + 1]); hi.y = f16x2_sat(v[i * 4 + 2], v[i * 4 + 3]);
        const __half2 h0 = *reinterpret_cast<const __half2*>(&hi.x), h1 = *reinterpret_cast<const __half2*>(&hi.y);
        // (a value beyond the fp16 range saturates in hi; its remainder saturates in lo: still finite)
        lo.x = f16x2_sat(v[i * 4] - __low2float(h0), v[i * 4 + 1] - __high2float(h0));
        lo.y = f16x2_sat(v[i * 4 + 2] - __low2float(h1), v[i * 4 + 3] - __high2float(h1));
      } else {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i * 4], v[i * 4 + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i * 4 + 2], v[i * 4 + 3]);
        __nv_bfloat162 l0 = __floats2bfloat162_rn(v[i * 4] - __low2float(h0), v[i * 4 + 1] - __high2float(h0));
        __nv_bfloat162 l1 = __floats2bfloat162_rn(v[i * 4 + 2] - __low2float(h1), v[i * 4 + 3] - __high2float(h1));
        hi.x = *reinterpret_cast<uint32_t*>(&h0); hi.y = *reinterpret_cast<uint32_t*>(&h1);
        lo.x = *reinterpret_cast<uint32_t*>(&l0); lo.y = *reinterpret_cast<uint32_t*>(&l1);
      }
      *reinterpret_cast<uint2*>(dst) = hi;
      *reinterpret_cast<uint2*>(dst + 64) = lo;
    }
  }
};


// LN(x) * g + b, then optionally the TimeBlock modulation * (1 + scale) + shift and SiLU (cross_attention.py:437-438);
// m points at [scale(D) | shift(D)] or is null.
// FAST (bf16 outputs only): SiLU through ex2.approx / rcp.approx -- about 2 ulp of float, far below the bf16 rounding
// that follows -- instead of expf and an IEEE division (6x the instructions of the LayerNorm itself).
template <int D, bool FAST = false>
__device__ __forceinline__ void ln_row_finish(RowVec<D>& r, const float* __restrict__ g, const float* __restrict__ b,
                                              const float* __restrict__ m, int lane) {
  r.normalize();
  r.affine(g, b, lane);
  if (m) {
#pragma unroll
    for (int i = 0; i < RowVec<D>::NV; ++i) {
      float4 sc = *reinterpret_cast<const float4*>(m + i * 128 + lane * 4);
      float4 sh = *reinterpret_cast<const float4*>(m + D + i * 128 + lane * 4);
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float y = r.v[i * 4 + j] * (1.0f + scv[j]) + shv[j];
        if constexpr (FAST) r.v[i * 4 + j] = __fdividef(y, 1.0f + __expf(-y));
        else r.v[i * 4 + j] = act_apply(y, CFB_ACT_SILU);
      }
    }
  }
}

}  // namespace cfb
