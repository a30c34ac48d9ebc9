// Library-level C ABI: error reporting, switches, and the unit-op entry points the tests use.
#include "common.cuh"
#include "kernels.cuh"
#include "rowblock.cuh"
#include <cstdlib>

namespace cfb {

static thread_local char g_err[1024] = "";
std::atomic<unsigned long long> g_launches{0};
thread_local bool t_capturing = false;
int g_use_pdl = getenv("CFB_PDL") ? atoi(getenv("CFB_PDL")) : 1;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int init_gemm_tc_kernels();
int init_attention_kernels();
int tc_trace_read(unsigned long long out[16]);
extern int g_rb_trace_kind, g_rb_trace_layer;   // denoiser.cu

static int ensure_device() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: convofusion_b200 has no CPU fallback");
    return CFB_ERR_NO_DEVICE;
  }
  CFB_TRY(init_gemm_tc_kernels());
  return init_attention_kernels();
}

}  // namespace cfb

using namespace cfb;

extern "C" {

// Debug only (not part of include/convofusion_b200.h): phase timestamps of the last traced tcgen05 GEMM CTA.
int cfb_debug_tc_trace(unsigned long long* out16) { return cfb::tc_trace_read(out16); }

// Debug only: fault record of the row-block kernel's guarded barrier waits {code, block, warp, stage, barrier, parity};
// host-mapped memory, readable after the context has been lost to the trap.
int cfb_debug_rb_fault(unsigned* out8) {
  const unsigned* f = cfb::rowblock_fault_record();
  for (int i = 0; i < 8; ++i) out8[i] = f ? f[i] : 0u;
  return CFB_OK;
}

// Debug only: arm phase tracing of the next launch of program `kind` of `layer` (eager launches only), read it back.
int cfb_debug_rb_trace_arm(int kind, int layer, int n) {
  cfb::g_rb_trace_kind = kind; cfb::g_rb_trace_layer = layer;
  cfb::rowblock_trace_arm(n);
  return CFB_OK;
}
int cfb_debug_rb_trace_read(unsigned long long* out64) { return cfb::rowblock_trace_read(out64); }

int cfb_abi_version(void) { return CFB_ABI_VERSION; }
const char* cfb_last_error(void) { return g_err; }
unsigned long long cfb_launch_count(void) { return g_launches.load(); }

int cfb_set_gemm_backend(int backend) {
  CFB_CHECK(backend >= CFB_GEMM_AUTO && backend <= CFB_GEMM_TCGEN05, "unknown gemm backend %d", backend);
  g_gemm_backend = backend;
  return CFB_OK;
}

int cfb_linear(const void* A, int a_bf16, const void* W, const float* bias, void* out, int out_bf16, int M, int N,
               int K, int act, int a_act, int accumulate, int backend, cfb_stream stream) {
  CFB_CHECK(A && W && out && M > 0 && N > 0 && K > 0, "cfb_linear: bad argument");
  CFB_TRY(ensure_device());
  Epilogue ep{};
  ep.bias = bias; ep.bias_period = 1; ep.act = act; ep.accumulate = accumulate; ep.out_bf16 = out_bf16;
  ep.out = out; ep.ldo = N; ep.replicate = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (backend == CFB_GEMM_SIMT) return gemm_simt(A, a_bf16, K, W, a_bf16, K, M, N, K, a_act, ep, st);
  if (backend == CFB_GEMM_TCGEN05) {
    CFB_CHECK(a_bf16 && !a_act, "cfb_linear: the tcgen05 engine takes bf16 operands and no input activation");
    return gemm_tc((const bf16*)A, K, (const bf16*)W, K, M, N, K, ep, st);
  }
  return gemm(A, a_bf16, K, W, a_bf16, K, M, N, K, a_act, ep, st);
}

int cfb_keypoints3d(const float* feats, long long rows, float* out, cfb_stream stream) {
  CFB_CHECK(feats && out && rows >= 0, "cfb_keypoints3d: bad argument");
  CFB_CHECK(feats != out, "cfb_keypoints3d: in-place use is not supported (finger joints read their wrist)");
  CFB_TRY(ensure_device());
  return keypoints3d(feats, rows, out, (cudaStream_t)stream);
}

int cfb_layernorm(const float* x, const float* g, const float* b, void* out, int out_bf16, int rows, int d,
                  cfb_stream stream) {
  CFB_CHECK(x && g && b && out, "cfb_layernorm: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (out_bf16) return ln_rows<bf16>(x, g, b, nullptr, nullptr, 0, (bf16*)out, rows, d, st);
  return ln_rows<float>(x, g, b, nullptr, nullptr, 0, (float*)out, rows, d, st);
}

int cfb_mha(const void* q, int ldq, const void* k, const void* v, int ldk, void* out, int ldo, int is_bf16, int n,
            int Lq, int Lk, int n_heads, int head_dim, const int32_t* kv_len, cfb_stream stream) {
  CFB_CHECK(q && k && v && out, "cfb_mha: null argument");
  CFB_TRY(ensure_device());
  cudaStream_t st = (cudaStream_t)stream;
  if (is_bf16)
    return mha<bf16>((const bf16*)q, ldq, (const bf16*)k, (const bf16*)v, ldk, (bf16*)out, ldo, n, Lq, Lk, n_heads,
                     head_dim, kv_len, st);
  return mha<float>((const float*)q, ldq, (const float*)k, (const float*)v, ldk, (float*)out, ldo, n, Lq, Lk, n_heads,
                    head_dim, kv_len, st);
}

int cfb_audio_encoder(const float* mel, int rows, const float* w0, const float* b0, const float* w1, const float* b1,
                      const float* w2, const float* b2, int n_mel, int hidden, int d_out, float* tmp0, float* tmp1,
                      float* out, cfb_stream stream) {
  // audioenc.py:13-34: Linear(80,256) -> LeakyReLU(0.1) -> Linear(256,512) -> LeakyReLU(0.1) -> Linear(512,512)
  CFB_CHECK(mel && w0 && w1 && w2 && tmp0 && tmp1 && out && rows > 0, "cfb_audio_encoder: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  Epilogue e0{}; e0.bias = b0; e0.bias_period = 1; e0.act = CFB_ACT_LEAKY01; e0.out = tmp0; e0.ldo = hidden; e0.replicate = 1;
  CFB_TRY(gemm_simt(mel, 0, n_mel, w0, 0, n_mel, rows, hidden, n_mel, 0, e0, st));
  Epilogue e1{}; e1.bias = b1; e1.bias_period = 1; e1.act = CFB_ACT_LEAKY01; e1.out = tmp1; e1.ldo = d_out; e1.replicate = 1;
  CFB_TRY(gemm_simt(tmp0, 0, hidden, w1, 0, hidden, rows, d_out, hidden, 0, e1, st));
  Epilogue e2{}; e2.bias = b2; e2.bias_period = 1; e2.out = out; e2.ldo = d_out; e2.replicate = 1;
  return gemm_simt(tmp1, 0, d_out, w2, 0, d_out, rows, d_out, d_out, 0, e2, st);
}

int cfb_guidance_sched_step(const float* eps, float* x, const float* noise, const float* coef_dev, int n_branch,
                            int n_clips, int n_per_clip, int kind, int clip_sample, float guidance_scale,
                            cfb_stream stream) {
  CFB_CHECK(eps && x && coef_dev, "cfb_guidance_sched_step: null argument");
  StepArgs a{};
  a.eps = eps; a.x = x; a.noise = noise; a.coef = coef_dev; a.n_branch = n_branch; a.n_clips = n_clips;
  a.full_last = n_branch == CFB_N_BRANCH; a.n_per_clip = n_per_clip; a.n_steps = 1; a.kind = kind; a.clip_sample = clip_sample; a.guidance_scale = guidance_scale;
  return guidance_sched_step(a, (cudaStream_t)stream);
}

}  // extern "C"
