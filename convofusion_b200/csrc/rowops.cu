// Row-wise memory-bound kernels: LayerNorm (+ TimeBlock modulation + SiLU) producing the next GEMM's
// A operand, conditioning-memory construction / per-step normalisation, timestep sinusoid, casts.
// One warp owns one row of D floats (D = 512 denoiser, 128 VAE): 128-bit loads, statistics by warp
// shuffle, two-pass variance (mean first, then sum of squared deviations) like ATen's LayerNorm.
#include <algorithm>
#include "common.cuh"
#include "kernels.cuh"
#include "rowvec.cuh"

namespace cfb {

namespace {

// out = LN(x) * g + b  [ * (1 + scale) + shift -> SiLU ]           (cross_attention.py:437-438)
// OUT: 1 = T, 2 = two bf16 terms [hi | lo], 3 = fp16 in the bf16 buffer, 4 = two fp16 terms [hi | lo]
template <typename T, int D, int OUT = 1>
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                      const float* __restrict__ b, const float* __restrict__ mod,
                                                      const int* __restrict__ step_ptr, long long mod_step_stride,
                                                      T* __restrict__ out, int rows) {
  pdl_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  RowVec<D> r;
  r.load(x + (size_t)row * D, lane);
  ln_row_finish<D, sizeof(T) == 2>(r, g, b, mod ? mod + (step_ptr ? (size_t)(*step_ptr) * mod_step_stride : 0) : nullptr, lane);
  if constexpr (OUT == 2 && sizeof(T) == 2) r.template store_split<false>(reinterpret_cast<bf16*>(out) + (size_t)row * 2 * D, lane);
  else if constexpr (OUT == 4 && sizeof(T) == 2) r.template store_split<true>(reinterpret_cast<bf16*>(out) + (size_t)row * 2 * D, lane);
  else if constexpr (OUT == 3 && sizeof(T) == 2) r.store_f16(reinterpret_cast<bf16*>(out) + (size_t)row * D, lane);
  else r.store(out + (size_t)row * D, lane);
}

__global__ void __launch_bounds__(256) bf16_to_f16_kernel(const bf16* __restrict__ in, __half* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
    out[i] = from_f32<__half>(__bfloat162float(in[i]));
}

// mem_c[row] = cond[row] + stream_emb[x] + pe[pos]          (denoiser.py:332-353, time-independent part)
struct MemBuildArgs {
  const float* cond[CFB_N_STREAMS];
  int row_base[CFB_N_STREAMS + 1];  // first global row of each stream
  int len[CFB_N_STREAMS];
};
template <int D>
__global__ void __launch_bounds__(256) mem_build_kernel(MemBuildArgs a, const float* __restrict__ stream_emb,
                                                        const float* __restrict__ pe, float* __restrict__ mem_c) {
  pdl_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= a.row_base[CFB_N_STREAMS]) return;
  int x = 0;
#pragma unroll
  for (int i = 1; i < CFB_N_STREAMS; ++i) x += (row >= a.row_base[i]);
  const int local = row - a.row_base[x];
  const int pos = local % a.len[x];
  RowVec<D> r;
  r.load(a.cond[x] + (size_t)local * D, lane);
  r.add(stream_emb + (size_t)x * D, lane);
  r.add(pe + (size_t)pos * D, lane);
  r.store(mem_c + (size_t)row * D, lane);
}

// mem_hat[row] = (mem_c[row] + temb - mean) * rstd : the affine-free LayerNorm every layer's
// {stream}_norm shares (denoiser.py:255-261 + cross_attention.py:581-585; gamma/beta live in w_qx/w_fu).
template <typename T, int D, bool F16 = false>
__global__ void __launch_bounds__(256) mem_hat_kernel(const float* __restrict__ mem_c, const float* __restrict__ temb,
                                                      const int* __restrict__ step_ptr, T* __restrict__ out, int rows) {
  pdl_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  RowVec<D> r;
  r.load(mem_c + (size_t)row * D, lane);
  r.add(temb + (step_ptr ? (size_t)(*step_ptr) * D : 0), lane);
  r.normalize();
  if constexpr (F16 && sizeof(T) == 2) r.store_f16(reinterpret_cast<bf16*>(out) + (size_t)row * D, lane);
  else r.store(out + (size_t)row * D, lane);
}

// embeddings.py:245-285 with flip_sin_to_cos=True, freq_shift=0: [cos(t f_k), sin(t f_k)], f_k = exp(-ln(1e4) k / half)
__global__ void time_sinusoid_kernel(const float* __restrict__ t, float* __restrict__ out, int n, int dim) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= n * half) return;
  const int row = i / half, k = i % half;
  const float exponent = (-9.210340371976184f * (float)k) / (float)half;  // -ln(10000) * k / half, float like torch
  const float e = t[row] * expf(exponent);
  out[(size_t)row * dim + k] = cosf(e);
  out[(size_t)row * dim + half + k] = sinf(e);
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ in, T* __restrict__ out, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = from_f32<T>(in[i]);
}

// out[r, 0:d] = a[r], out[r, d:2d] = b[r]                   (cross_attention.py:114 torch.cat([x, xs.pop()]))
template <typename T>
__global__ void concat2_kernel(const float* __restrict__ a, const float* __restrict__ b, T* __restrict__ out,
                               int rows, int d) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * 2 * d) return;
  const int r = (int)(i / (2 * d)), c = (int)(i % (2 * d));
  out[i] = from_f32<T>(c < d ? a[(size_t)r * d + c] : b[(size_t)r * d + c - d]);
}

// rows [B*L, D]: out[b, l] = pe[l] (+ src[b, l])           (vae.py:277,321-322)
template <typename T>
__global__ void add_pe_kernel(const float* __restrict__ src, const float* __restrict__ pe, T* __restrict__ out,
                              int n_batch, int L, int d) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_batch * L * d) return;
  const int c = (int)(i % d);
  const int l = (int)((i / d) % L);
  const float v = pe[(size_t)l * d + c] + (src ? src[i] : 0.f);
  out[i] = from_f32<T>(v);
}

// vae.py:362: zero frames at or beyond each clip's length.
__global__ void mask_frames_kernel(float* __restrict__ out, const int* __restrict__ lengths, int n_batch, int L, int d) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_batch * L * d) return;
  const int l = (int)((i / d) % L), b = (int)(i / ((long long)d * L));
  if (l >= lengths[b]) out[i] = 0.f;
}

// vae.py:176-186: per 16-frame chunk, subtract the first frame's root x and z from every frame of the chunk.
__global__ void chunk_root_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int nf, int chunk) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / nf;
  const int c = (int)(i % nf);
  float v = in[i];
  if (c < 3) v = __fsub_rn(v, __fmul_rn(in[(r - r % chunk) * nf + c], c == 1 ? 0.f : 1.f));
  out[i] = v;
}

// vae.py:204-224: token rows of one part, sample-major: [n_tok distribution tokens ; chunk frame embeddings] + pe.
__global__ void enc_assemble_kernel(const float* __restrict__ emb, const float* __restrict__ tokens,
                                    const float* __restrict__ pe, float* __restrict__ h, int n, int n_tok, int chunk, int d) {
  pdl_sync();
  const int L = n_tok + chunk;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * L * d) return;
  const int c = (int)(i % d);
  const int l = (int)((i / d) % L);
  const long long s = i / ((long long)d * L);
  const float v = l < n_tok ? tokens[(size_t)l * d + c] : emb[((size_t)s * chunk + (l - n_tok)) * d + c];
  h[i] = v + pe[(size_t)l * d + c];
}

// vae.py:250-260: mu = token 0, logvar = token 1 of the normalised encoder output; std = exp(logvar)^0.5.
__global__ void enc_dist_kernel(const float* __restrict__ y, float* __restrict__ mu, float* __restrict__ sd, int n, int L, int d) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * d) return;
  const long long s = i / d;
  const int c = (int)(i % d);
  mu[i] = y[(s * L) * d + c];
  sd[i] = sqrtf(expf(y[(s * L + 1) * d + c]));
}

// base.py:204-209: p = f / 3; p[43:] += p[11]; p[23:43] += p[7]; p[1:] += p[0]   (in that order, fp32)
__global__ void keypoints3d_kernel(const float* __restrict__ f, float* __restrict__ out, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long row = i / 189;
  const int c = (int)(i % 189), j = c / 3, a = c % 3;
  const float* fr = f + row * 189;
  float v = __fdiv_rn(fr[c], 3.0f);
  if (j >= 43) v = __fadd_rn(v, __fdiv_rn(fr[11 * 3 + a], 3.0f));
  else if (j >= 23) v = __fadd_rn(v, __fdiv_rn(fr[7 * 3 + a], 3.0f));
  if (j >= 1) v = __fadd_rn(v, __fdiv_rn(fr[a], 3.0f));
  out[i] = v;
}

}  // namespace

int keypoints3d(const float* feats, long long rows, float* out, cudaStream_t st) {
  const long long n = rows * 189;
  if (n <= 0) return CFB_OK;
  launch_k(keypoints3d_kernel, (unsigned)((n + 255) / 256), 256, 0, st, feats, out, n);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int chunk_root(const float* in, float* out, long long n_rows, int nf, int chunk, cudaStream_t st) {
  const long long n = n_rows * nf;
  launch_k(chunk_root_kernel, (unsigned)((n + 255) / 256), 256, 0, st, in, out, n, nf, chunk);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
int enc_assemble(const float* emb, const float* tokens, const float* pe, float* h, int n, int n_tok, int chunk, int d,
                 cudaStream_t st) {
  const long long t = (long long)n * (n_tok + chunk) * d;
  launch_k(enc_assemble_kernel, (unsigned)((t + 255) / 256), 256, 0, st, emb, tokens, pe, h, n, n_tok, chunk, d);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
int enc_dist(const float* y, float* mu, float* sd, int n, int L, int d, cudaStream_t st) {
  const long long t = (long long)n * d;
  launch_k(enc_dist_kernel, (unsigned)((t + 255) / 256), 256, 0, st, y, mu, sd, n, L, d);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int bf16_to_f16(const bf16* in, bf16* out, size_t n, cudaStream_t st) {
  if (n == 0) return CFB_OK;
  bf16_to_f16_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(in, reinterpret_cast<__half*>(out), n);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

template <typename T>
int ln_rows(const float* x, const float* g, const float* b, const float* mod, const int* step_ptr,
            long long mod_step_stride, T* out, int rows, int d, cudaStream_t st, int terms) {
  if (rows <= 0 || debug_skip(1)) return CFB_OK;
  dim3 grid(ceil_div(rows, 8));
  if (terms == 2 || terms == 4) {      // [hi | lo] per 64 columns, row stride 2 d (16-bit outputs of the denoiser only); 4: fp16 terms
    CFB_CHECK(sizeof(T) == 2 && d == 512, "ln_rows: two-term output needs a 16-bit buffer and d = 512");
    if (terms == 2) launch_k(ln_rows_kernel<T, 512, 2>, grid, 256, 0, st, x, g, b, mod, step_ptr, mod_step_stride, out, rows);
    else launch_k(ln_rows_kernel<T, 512, 4>, grid, 256, 0, st, x, g, b, mod, step_ptr, mod_step_stride, out, rows);
    CFB_LAUNCH_CHECK();
    return CFB_OK;
  }
  if (terms == 3) {      // fp16 values in the bf16 buffer (same layout)
    CFB_CHECK(sizeof(T) == 2 && (d == 512 || d == 128), "ln_rows: fp16 output needs a 16-bit buffer and d = 512 or 128");
    if (d == 512) launch_k(ln_rows_kernel<T, 512, 3>, grid, 256, 0, st, x, g, b, mod, step_ptr, mod_step_stride, out, rows);
    else launch_k(ln_rows_kernel<T, 128, 3>, grid, 256, 0, st, x, g, b, mod, step_ptr, mod_step_stride, out, rows);
    CFB_LAUNCH_CHECK();
    return CFB_OK;
  }
  if (d == 512) launch_k(ln_rows_kernel<T, 512>, grid, 256, 0, st, x, g, b, mod, step_ptr, mod_step_stride, out, rows);
  else if (d == 128) launch_k(ln_rows_kernel<T, 128>, grid, 256, 0, st, x, g, b, mod, step_ptr, mod_step_stride, out, rows);
  else { set_error("ln_rows: d_model %d unsupported (128 or 512)", d); return CFB_ERR_INVALID; }
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int ln_rows<float>(const float*, const float*, const float*, const float*, const int*, long long, float*, int, int, cudaStream_t, int);
template int ln_rows<bf16>(const float*, const float*, const float*, const float*, const int*, long long, bf16*, int, int, cudaStream_t, int);

int mem_build(const float* const cond[CFB_N_STREAMS], const int n_slots[CFB_N_STREAMS], const int len[CFB_N_STREAMS],
              const float* stream_emb, const float* pe, float* mem_c, int d, cudaStream_t st) {
  CFB_CHECK(d == 512, "mem_build: d_model %d unsupported", d);
  MemBuildArgs a;
  int base = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    a.cond[x] = cond[x]; a.row_base[x] = base; a.len[x] = len[x] > 0 ? len[x] : 1;
    base += n_slots[x] * len[x];
  }
  a.row_base[CFB_N_STREAMS] = base;
  if (base == 0) return CFB_OK;
  launch_k(mem_build_kernel<512>, ceil_div(base, 8), 256, 0, st, a, stream_emb, pe, mem_c);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

template <typename T>
int mem_hat(const float* mem_c, const float* temb, const int* step_ptr, T* out, int rows, int d, cudaStream_t st, int f16) {
  CFB_CHECK(d == 512, "mem_hat: d_model %d unsupported", d);
  if (rows <= 0) return CFB_OK;
  if (f16 && sizeof(T) == 2) launch_k(mem_hat_kernel<T, 512, true>, ceil_div(rows, 8), 256, 0, st, mem_c, temb, step_ptr, out, rows);
  else launch_k(mem_hat_kernel<T, 512>, ceil_div(rows, 8), 256, 0, st, mem_c, temb, step_ptr, out, rows);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int mem_hat<float>(const float*, const float*, const int*, float*, int, int, cudaStream_t, int);
template int mem_hat<bf16>(const float*, const float*, const int*, bf16*, int, int, cudaStream_t, int);

int time_sinusoid(const float* t, float* out, int n, int dim, cudaStream_t st) {
  const int total = n * (dim / 2);
  launch_k(time_sinusoid_kernel, ceil_div(total, 256), 256, 0, st, t, out, n, dim);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

template <typename T>
int cast_rows(const float* in, T* out, long long n, cudaStream_t st) {
  if (n <= 0) return CFB_OK;
  launch_k(cast_kernel<T>, (unsigned)((n + 255) / 256), 256, 0, st, in, out, n);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int cast_rows<float>(const float*, float*, long long, cudaStream_t);
template int cast_rows<bf16>(const float*, bf16*, long long, cudaStream_t);
template int cast_rows<__half>(const float*, __half*, long long, cudaStream_t);

template <typename T>
int concat2(const float* a, const float* b, T* out, int rows, int d, cudaStream_t st) {
  const long long n = (long long)rows * 2 * d;
  launch_k(concat2_kernel<T>, (unsigned)((n + 255) / 256), 256, 0, st, a, b, out, rows, d);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int concat2<float>(const float*, const float*, float*, int, int, cudaStream_t);
template int concat2<bf16>(const float*, const float*, bf16*, int, int, cudaStream_t);
template int concat2<__half>(const float*, const float*, __half*, int, int, cudaStream_t);

template <typename T>
int add_pe(const float* src, const float* pe, T* out, int n_batch, int L, int d, cudaStream_t st) {
  const long long n = (long long)n_batch * L * d;
  launch_k(add_pe_kernel<T>, (unsigned)((n + 255) / 256), 256, 0, st, src, pe, out, n_batch, L, d);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int add_pe<float>(const float*, const float*, float*, int, int, int, cudaStream_t);
template int add_pe<bf16>(const float*, const float*, bf16*, int, int, int, cudaStream_t);
template int add_pe<__half>(const float*, const float*, __half*, int, int, int, cudaStream_t);

int mask_frames(float* out, const int* lengths, int n_batch, int L, int d, cudaStream_t st) {
  const long long n = (long long)n_batch * L * d;
  launch_k(mask_frames_kernel, (unsigned)((n + 255) / 256), 256, 0, st, out, lengths, n_batch, L, d);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

}  // namespace cfb
