// Shared helpers for the convofusion_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <atomic>
#include "../../include/convofusion_b200.h"

namespace cfb {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;   // kernels launched (graph replays included)
extern thread_local bool t_capturing;                    // launches recorded into a graph are counted per replay

#define CFB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      cfb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CFB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define CFB_CHECK(cond, ...)                                                             \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      cfb::set_error(__VA_ARGS__);                                                       \
      return CFB_ERR_INVALID;                                                            \
    }                                                                                    \
  } while (0)

#define CFB_TRY(expr)                                                                    \
  do {                                                                                   \
    int _s = (expr);                                                                     \
    if (_s != CFB_OK) return _s;                                                         \
  } while (0)

// Every kernel launch goes through this so gpu_launches is an honest count.
#define CFB_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    if (!cfb::t_capturing) ++cfb::g_launches;                                            \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      cfb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return CFB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Diagnosis only (build with -DCFB_DEBUG_SKIP): env CFB_SKIP is a bit mask of kernel classes whose launches are
// dropped (1 LayerNorm rows, 2 self-attention / MHA, 4 shared softmax, 8 per-pair cross-attention, 16 tcgen05 GEMM,
// 32 grouped GEMM) to measure what each class costs inside the concurrent step.  Results are garbage.
#ifdef CFB_DEBUG_SKIP
#include <cstdlib>
static inline bool debug_skip(int bit) {
  static int mask = -1;
  if (mask < 0) { const char* e = getenv("CFB_SKIP"); mask = e ? atoi(e) : 0; }
  return (mask & bit) != 0;
}
#else
static inline bool debug_skip(int) { return false; }
#endif

// Programmatic dependent launch.  Every kernel of this library starts with pdl_sync() (tcgen05 GEMM: after its
// shared-memory/TMEM prologue): `launch_dependents` lets the NEXT kernel's CTAs be scheduled as soon as all of this
// kernel's CTAs have started, `wait` blocks until the PREVIOUS kernel has completed and its writes are visible.
// With ~190 kernels of 3-10 us per sampling step this hides launch latency and prologues behind predecessors' tails.
extern int g_use_pdl;
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
// float -> fp16 with saturation to +-65504 in the conversion itself (cvt.rn.satfinite: one instruction, no min / max)
__device__ __forceinline__ uint32_t f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}
template <> __device__ __forceinline__ __half from_f32<__half>(float v) {
  const uint32_t r = f16x2_sat(v, 0.f);
  return __ushort_as_half((unsigned short)(r & 0xffffu));
}
// Two floats -> one 32-bit pair of 16-bit values: bf16, or fp16 (clamped to its range) when f16 is set.
__device__ __forceinline__ uint32_t pack16(float a, float b, int f16) {
  if (f16) return f16x2_sat(a, b);
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case CFB_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));  // exact-erf GELU (F.gelu default)
    case CFB_ACT_SILU: return v / (1.0f + expf(-v));
    case CFB_ACT_RELU: return v > 0.f ? v : 0.f;
    case CFB_ACT_LEAKY01: return v > 0.f ? v : 0.1f * v;
    default: return v;
  }
}

// Activations for values that are rounded to bf16 right afterwards: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7,
// one ex2.approx and one rcp.approx instead of the ~25-instruction erff), SiLU through ex2.approx / rcp.approx.
__device__ __forceinline__ float act_apply_fast(float v, int act) {
  switch (act) {
    case CFB_ACT_GELU: {
      const float z = fabsf(v) * 0.70710678118654752440f;
      const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
      float p = fmaf(1.061405429f, t, -1.453152027f);
      p = fmaf(p, t, 1.421413741f);
      p = fmaf(p, t, -0.284496736f);
      p = fmaf(p, t, 0.254829592f);
      const float e = 1.0f - p * t * __expf(-z * z);          // erf(|v| / sqrt 2)
      return 0.5f * v * (1.0f + copysignf(e, v));
    }
    case CFB_ACT_SILU: return __fdividef(v, 1.0f + __expf(-v));
    default: return act_apply(v, act);
  }
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

// ---- fp32-accurate tensor-core GEMMs (gemm_split.cu): resources of one call chain
struct SplitCache;
struct SplitCtx {
  void* a_ws; size_t a_ws_bytes;   // stream-ordered scratch for the split of a dynamic A operand
  void* w_ws; size_t w_ws_bytes;   // ... of a dynamic W operand
  SplitCache* cache;               // splits of static operands (weights), filled on first use outside stream capture
  int scheme;                      // 0: 3 x 3 terms (fp32-accurate), 1: A hi+lo x W bf16, 2: bf16 x bf16 (gemm_split.cu)
};
SplitCache* split_cache_create();
void split_cache_destroy(SplitCache* c);
size_t split_cache_bytes(const SplitCache* c);

// ---- epilogue description shared by the SIMT and tcgen05 GEMMs -------------------
struct Epilogue {
  const float* bias;   // [bias_period, N] (period 1 = ordinary bias) or nullptr
  int bias_period;     // rows r use bias[(r % period) * N + n]
  int act;             // cfb_act applied after bias
  int accumulate;      // out += value (float out only): residual update
  int out_bf16;        // element type of out
  void* out;
  int ldo;
  int replicate;       // write `replicate` copies, copy c at out + c * rep_stride elements
  long long rep_stride;
  int tma_out;               // set by gemm_tc: output tile leaves through TMA store / reduce-add (internal)
  // fp32 operands on the tensor cores (3-way bf16 split, gemm_split.cu): null = CUDA-core GEMM for fp32 operands
  const SplitCtx* split;
  int a_static, w_static;    // the operand is a weight (its split is cached) rather than an activation
  int a_from_ln;             // precision study (scheme 3): the A operand is a LayerNorm output
  int a_terms;               // bf16 A operand with 2 terms per value ([hi | lo] per 64 columns, lda >= 2 K); 0 / 1 = plain
  int out_f16;               // out_bf16 outputs are written as fp16 (clamped) instead of bf16
  int ab_f16;                // both 16-bit operands hold fp16, not bf16, values (tcgen05 path only; kind::f16 does not mix)
};

// GEMM entry points (gemm_simt.cu / gemm_tc.cu). A [M,K] and W [N,K] are K-contiguous.
int gemm_simt(const void* A, int a_bf16, int lda, const void* W, int w_bf16, int ldw, int M, int N, int K,
              int a_act, const Epilogue& ep, cudaStream_t st);
bool gemm_tc_supported(int M, int N, int K, int lda, int ldw);
// w_rows (0 = N): number of valid rows of W; rows in [w_rows, N) read as zeros (TMA out-of-bounds fill).
int gemm_tc(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const Epilogue& ep,
            cudaStream_t st, int w_rows = 0);
// Grouped tcgen05 GEMM: group z computes rows [row_start, row_start+rows) of A_z . W_z^T (+bias_z) into out_z;
// A_z / out_z are base pointers indexed by ABSOLUTE row; N, K, lda, ldw, ep flags are shared.
constexpr int TC_MAX_GROUPS = 10;
struct TcGroup {
  const bf16* A; const bf16* W; const float* bias; void* out; int row_start; int rows;
};
int gemm_tc_grouped(const TcGroup* groups, int n_groups, int a_rows_total, int lda, int ldw, int N, int K,
                    const Epilogue& ep, cudaStream_t st);
extern int g_gemm_backend;
bool gemm_split_supported(int M, int N, int K, long long lda, long long ldw);
int gemm_split(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K, const Epilogue& ep,
               cudaStream_t st);

// y = act(A W^T + b) dispatch: bf16 operands use tcgen05 when allowed, everything else SIMT.
int gemm(const void* A, int a_bf16, int lda, const void* W, int w_bf16, int ldw, int M, int N, int K,
         int a_act, const Epilogue& ep, cudaStream_t st);

}  // namespace cfb
