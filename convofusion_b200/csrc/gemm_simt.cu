// CUDA-core GEMM: y[M,N] (+)= act(a_act(A)[M,K] . W[N,K]^T + bias).
//
// Roles: (1) the float32 engine of the fp32 parity mode, (2) the engine for shapes the tcgen05
// kernel does not take (K not a multiple of 64, tiny row counts such as the per-schedule time
// tables), (3) the on-device cross-check for gemm_tc.cu in tests.  64x64x16 tiles, 256 threads,
// 4x4 outputs per thread, operands staged transposed in shared memory as float.
#include <type_traits>
#include "common.cuh"

namespace cfb {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <typename T>
__device__ __forceinline__ void load4(const T* __restrict__ base, int ld, int row, int nrows, int k, int K,
                                      bool vec_ok, float out[4]) {
  if (row < nrows) {
    const T* p = base + (size_t)row * ld + k;
    if (vec_ok && k + 3 < K) {
      if constexpr (sizeof(T) == 4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
      } else if constexpr (std::is_same<T, __half>::value) {
        uint2 v = *reinterpret_cast<const uint2*>(p);
        const float2 a = __half22float2(*reinterpret_cast<__half2*>(&v.x));
        const float2 b = __half22float2(*reinterpret_cast<__half2*>(&v.y));
        out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
      } else {
        uint2 v = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
        out[0] = __low2float(a); out[1] = __high2float(a); out[2] = __low2float(b); out[3] = __high2float(b);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = (k + j < K) ? to_f32<T>(p[j]) : 0.f;
    }
  } else {
    out[0] = out[1] = out[2] = out[3] = 0.f;
  }
}

template <typename TA, typename TW>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, int lda,
                                                        const TW* __restrict__ W, int ldw, int M, int N,
                                                        int K, int a_act, bool vec_a, bool vec_w, Epilogue ep) {
  pdl_sync();
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Ws[BK][BN + PAD];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = t >> 2, lk = (t & 3) * 4;  // loader: row within tile, k offset
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    float a[4], w[4];
    load4<TA>(A, lda, m0 + lr, M, k0 + lk, K, vec_a, a);
    load4<TW>(W, ldw, n0 + lr, N, k0 + lk, K, vec_w, w);
    if (a_act) {
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = act_apply(a[j], a_act);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) { As[lk + j][lr] = a[j]; Ws[lk + j][lr] = w[j]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
    const float* brow = ep.bias ? ep.bias + (size_t)(r % ep.bias_period) * N : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (brow ? brow[n] : 0.f);
      v = act_apply(v, ep.act);
      for (int c = 0; c < ep.replicate; ++c) {
        const size_t off = (size_t)c * ep.rep_stride + (size_t)r * ep.ldo + n;
        if (ep.out_bf16) {
          if (ep.out_f16) reinterpret_cast<__half*>(ep.out)[off] = from_f32<__half>(v);
          else reinterpret_cast<bf16*>(ep.out)[off] = __float2bfloat16_rn(v);
        } else {
          float* o = reinterpret_cast<float*>(ep.out) + off;
          *o = ep.accumulate ? (*o + v) : v;
        }
      }
    }
  }
}

}  // namespace

int gemm_simt(const void* A, int a_bf16, int lda, const void* W, int w_bf16, int ldw, int M, int N, int K,
              int a_act, const Epilogue& ep_in, cudaStream_t st) {
  CFB_CHECK(M > 0 && N > 0 && K > 0, "gemm_simt: empty problem %dx%dx%d", M, N, K);
  Epilogue ep = ep_in;
  if (ep.replicate < 1) ep.replicate = 1;
  if (ep.bias_period < 1) ep.bias_period = 1;
  CFB_CHECK(!(ep.accumulate && ep.out_bf16), "gemm: accumulate needs float output");
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  const int ea = a_bf16 ? 2 : 4, ew = w_bf16 ? 2 : 4;
  const bool vec_a = ((uintptr_t)A % (4 * ea) == 0) && (lda % 4 == 0);
  const bool vec_w = ((uintptr_t)W % (4 * ew) == 0) && (ldw % 4 == 0);
  CFB_CHECK(!ep.ab_f16 || (a_bf16 && w_bf16), "gemm_simt: fp16 operands must both be 16-bit");
  if (ep.ab_f16)
    launch_k(gemm_simt_kernel<__half, __half>, grid, 256, 0, st, (const __half*)A, lda, (const __half*)W, ldw, M, N, K, a_act, vec_a, vec_w, ep);
  else if (a_bf16 && w_bf16)
    launch_k(gemm_simt_kernel<bf16, bf16>, grid, 256, 0, st, (const bf16*)A, lda, (const bf16*)W, ldw, M, N, K, a_act, vec_a, vec_w, ep);
  else if (!a_bf16 && !w_bf16)
    launch_k(gemm_simt_kernel<float, float>, grid, 256, 0, st, (const float*)A, lda, (const float*)W, ldw, M, N, K, a_act, vec_a, vec_w, ep);
  else if (a_bf16 && !w_bf16)
    launch_k(gemm_simt_kernel<bf16, float>, grid, 256, 0, st, (const bf16*)A, lda, (const float*)W, ldw, M, N, K, a_act, vec_a, vec_w, ep);
  else
    launch_k(gemm_simt_kernel<float, bf16>, grid, 256, 0, st, (const float*)A, lda, (const bf16*)W, ldw, M, N, K, a_act, vec_a, vec_w, ep);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int g_gemm_backend = CFB_GEMM_AUTO;

int gemm(const void* A, int a_bf16, int lda, const void* W, int w_bf16, int ldw, int M, int N, int K,
         int a_act, const Epilogue& ep, cudaStream_t st) {
  const bool tc_ok = a_bf16 && w_bf16 && !a_act && gemm_tc_supported(M, N, K, lda, ldw);
  if (g_gemm_backend == CFB_GEMM_TCGEN05)
    CFB_CHECK(tc_ok, "gemm: tcgen05 backend forced but shape %dx%dx%d (bf16=%d/%d) unsupported", M, N, K, a_bf16, w_bf16);
  if (tc_ok && g_gemm_backend != CFB_GEMM_SIMT)
    return gemm_tc((const bf16*)A, lda, (const bf16*)W, ldw, M, N, K, ep, st);
  CFB_CHECK(ep.a_terms != 2, "gemm: a two-term A operand needs the tcgen05 path (%dx%dx%d)", M, N, K);
  // fp32 operands: three-way bf16 split on the tensor cores when the caller supplies the resources for it
  if (!a_bf16 && !w_bf16 && !a_act && ep.split != nullptr && !ep.out_bf16 && g_gemm_backend != CFB_GEMM_SIMT &&
      gemm_split_supported(M, N, K, lda, ldw))
    return gemm_split((const float*)A, lda, (const float*)W, ldw, M, N, K, ep, st);
  return gemm_simt(A, a_bf16, lda, W, w_bf16, ldw, M, N, K, a_act, ep, st);
}

}  // namespace cfb
