// Attention kernels.
//  * mha_kernel: the core of torch.nn.MultiheadAttention on already-projected q/k/v (denoiser
//    self-attention 16x16 / 4 heads x 128; VAE self-attention 128x128 and cross-attention 128x8,
//    2 heads x 64) with an optional per-sample key length (key_padding_mask of lengths_to_mask).
//  * cross_kernel: the denoiser's five single-head cross-attentions in folded form: the query side
//    already carries W_k (qx = A_x norm2(tgt) + a_x), keys AND values are the affine-free
//    normalised memory rows, so one block per (batch entry, stream) computes
//    softmax(qx . mem_hat^T) . mem_hat over head_dim = d_model = 512.
#include "common.cuh"
#include "kernels.cuh"

namespace cfb {

namespace {

template <typename T>
__global__ void __launch_bounds__(128) mha_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k,
                                                  const T* __restrict__ v, int ldk, T* __restrict__ out, int ldo,
                                                  int Lq, int Lk, int hd, const int* __restrict__ kv_len,
                                                  int q_per_block) {
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
    } else {
      uint2 t = *reinterpret_cast<const uint2*>(row + i * 128 + lane * 4);
      __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
      __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
      v[i * 4] = __low2float(a); v[i * 4 + 1] = __high2float(a);
      v[i * 4 + 2] = __low2float(b); v[i * 4 + 3] = __high2float(b);
    }
  }
}

// grid (n_batch, n_streams); 256 threads.  qx/u rows are [n_batch * 16, 5 * 512].
template <typename T>
__global__ void __launch_bounds__(256) cross_kernel(const T* __restrict__ qx, const T* __restrict__ mem_hat,
                                                    T* __restrict__ u, CrossArgs a, int n_tokens) {
  extern __shared__ float sm[];
  const int bs = blockIdx.x, x = blockIdx.y;
  const int M = a.len[x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = CFB_N_STREAMS * CROSS_D;
  float* Qs = sm;                          // [16][512]
  float* Ss = Qs + CROSS_MAXQ * CROSS_D;   // [16][Mp]
  const int Mp = (M + 3) & ~3;
  const int slot = a.slot[x] ? a.slot[x][bs] : bs;
  const T* mem = mem_hat + ((size_t)a.row_base[x] + (size_t)slot * M) * CROSS_D;
  const uint8_t* msk = a.mask[x] ? a.mask[x] + (size_t)slot * M : nullptr;

  for (int i = threadIdx.x; i < n_tokens * CROSS_D; i += blockDim.x) {
    const int qi = i / CROSS_D, c = i % CROSS_D;
    Qs[i] = to_f32<T>(qx[(size_t)(bs * n_tokens + qi) * ld + x * CROSS_D + c]);
  }
  __syncthreads();
  // ---- scores: each warp walks keys j = warp, warp+8, ...; lanes split the 512 columns.
  for (int j = warp; j < M; j += 8) {
    float kv[16];
    load_row16<T>(mem + (size_t)j * CROSS_D, lane, kv);
    const bool masked = msk && msk[j];
    for (int qi = 0; qi < n_tokens; ++qi) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 qv = *reinterpret_cast<const float4*>(Qs + qi * CROSS_D + i * 128 + lane * 4);
        s = fmaf(qv.x, kv[i * 4], s); s = fmaf(qv.y, kv[i * 4 + 1], s);
        s = fmaf(qv.z, kv[i * 4 + 2], s); s = fmaf(qv.w, kv[i * 4 + 3], s);
      }
      s = warp_sum(s);
      if (lane == 0) Ss[qi * Mp + j] = masked ? -INFINITY : s;
    }
  }
  __syncthreads();
  // ---- softmax per query row (warp w: rows w, w+8)
  for (int qi = warp; qi < n_tokens; qi += 8) {
    float* srow = Ss + qi * Mp;
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
    for (int j = lane; j < M; j += 32) {
      const float p = srow[j] * inv;
      srow[j] = p;
      if (arow) arow[j] = p;
    }
  }
  __syncthreads();
  // ---- u = P . mem_hat: thread owns columns {2t, 2t+1}
  const int c = threadIdx.x * 2;
  float acc[CROSS_MAXQ][2];
#pragma unroll
  for (int qi = 0; qi < CROSS_MAXQ; ++qi) acc[qi][0] = acc[qi][1] = 0.f;
  for (int j = 0; j < M; ++j) {
    float v0, v1;
    if constexpr (sizeof(T) == 4) {
      const float2 t = *reinterpret_cast<const float2*>(mem + (size_t)j * CROSS_D + c);
      v0 = t.x; v1 = t.y;
    } else {
      const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(mem + (size_t)j * CROSS_D + c);
      v0 = __low2float(t); v1 = __high2float(t);
    }
#pragma unroll
    for (int qi = 0; qi < CROSS_MAXQ; ++qi) {
      const float p = Ss[qi * Mp + j];   // rows >= n_tokens are never stored below
      acc[qi][0] = fmaf(p, v0, acc[qi][0]);
      acc[qi][1] = fmaf(p, v1, acc[qi][1]);
    }
  }
  for (int qi = 0; qi < n_tokens; ++qi) {
    T* o = u + (size_t)(bs * n_tokens + qi) * ld + x * CROSS_D + c;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float2*>(o) = make_float2(acc[qi][0], acc[qi][1]);
    } else {
      *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(acc[qi][0], acc[qi][1]);
    }
  }
}

constexpr int ATT_MAX_SMEM = 160 * 1024;

}  // namespace

// Opt in to large dynamic shared memory once, outside any stream capture.
int init_attention_kernels() {
  static bool done = false;
  if (done) return CFB_OK;
  CFB_CUDA(cudaFuncSetAttribute(mha_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(mha_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(cross_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  CFB_CUDA(cudaFuncSetAttribute(cross_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_MAX_SMEM));
  done = true;
  return CFB_OK;
}

template <typename T>
int mha(const T* q, int ldq, const T* k, const T* v, int ldk, T* out, int ldo, int n, int Lq, int Lk, int n_heads,
        int head_dim, const int* kv_len, cudaStream_t st) {
  if (n <= 0) return CFB_OK;
  CFB_CHECK(Lq > 0 && Lk > 0 && head_dim > 0 && n_heads > 0, "mha: bad shape");
  const size_t smem = ((size_t)Lk * (head_dim + 1) + (size_t)Lk * head_dim + 4 * head_dim + 4 * Lk) * sizeof(float);
  CFB_CHECK(smem <= (size_t)ATT_MAX_SMEM, "mha: Lk=%d head_dim=%d needs %zu B of shared memory", Lk, head_dim, smem);
  const int q_per_block = Lq >= 64 ? 32 : Lq;
  dim3 grid(n, n_heads, ceil_div(Lq, q_per_block));
  mha_kernel<T><<<grid, 128, smem, st>>>(q, ldq, k, v, ldk, out, ldo, Lq, Lk, head_dim, kv_len, q_per_block);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int mha<float>(const float*, int, const float*, const float*, int, float*, int, int, int, int, int, int, const int*, cudaStream_t);
template int mha<bf16>(const bf16*, int, const bf16*, const bf16*, int, bf16*, int, int, int, int, int, int, const int*, cudaStream_t);

template <typename T>
int cross_attention(const T* qx, const T* mem_hat, T* u, const CrossArgs& a, int n_batch, int n_tokens, int d,
                    cudaStream_t st) {
  if (n_batch <= 0) return CFB_OK;
  CFB_CHECK(d == CROSS_D && n_tokens <= CROSS_MAXQ, "cross_attention: d=%d n_tokens=%d unsupported", d, n_tokens);
  int maxM = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    CFB_CHECK(a.len[x] > 0, "cross_attention: stream %d has no memory tokens", x);
    if (a.len[x] > maxM) maxM = a.len[x];
  }
  const size_t smem = ((size_t)CROSS_MAXQ * CROSS_D + (size_t)CROSS_MAXQ * ((maxM + 3) & ~3)) * sizeof(float);
  CFB_CHECK(smem <= (size_t)ATT_MAX_SMEM, "cross_attention: %d memory tokens exceed the shared-memory budget", maxM);
  dim3 grid(n_batch, CFB_N_STREAMS);
  cross_kernel<T><<<grid, 256, smem, st>>>(qx, mem_hat, u, a, n_tokens);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}
template int cross_attention<float>(const float*, const float*, float*, const CrossArgs&, int, int, int, cudaStream_t);
template int cross_attention<bf16>(const bf16*, const bf16*, bf16*, const CrossArgs&, int, int, int, cudaStream_t);

}  // namespace cfb
