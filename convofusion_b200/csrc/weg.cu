// Query-side backward of the denoiser for word-excitation guidance (convofusion.py:437-496, 298-388;
// tools/word_excitation_guidance.py:55-62): the gradient of a loss on the text-stream attention maps with respect to
// the latents.  Only input gradients are needed (no weight gradients), the batch is the text-only branch of ONE clip
// (the reference asserts batch size 1), i.e. 16 query rows, so these are small fp32 CUDA-core kernels: every launch
// is bound by reading the layer's weights once (21 MB per layer), not by arithmetic.  The orchestration (forward with
// saved activations, then the reverse walk over the layers) lives in denoiser.cu (cfb_denoiser_weg_forward / _backward).
#include "common.cuh"
#include "kernels.cuh"

namespace cfb {

namespace {

// dX[M, K] (+)= dY[M, N] . W[N, K]   (backward of y = x W^T for the input; W row-major [out = N, in = K]).
// grid (K / 32, ceil(M / 16)); 8 warps split N (warp w takes n = w, w + 8, ...), lanes take 32 consecutive k: every
// weight row segment is one coalesced 128-byte read and W is read exactly once per 16-row block.  The eight partial
// sums are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256) linear_bwd_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ W,
                                                         int ldw, float* __restrict__ dX, int ldx, int M, int N, int K,
                                                         int accumulate) {
  pdl_sync();
  __shared__ float red[8][16][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 32 + lane, m0 = blockIdx.y * 16;
  const int mrows = M - m0 < 16 ? M - m0 : 16;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  if (k < K) {
    for (int n = warp; n < N; n += 8) {
      const float w = __ldg(W + (size_t)n * ldw + k);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < mrows) acc[i] = fmaf(__ldg(dY + (size_t)(m0 + i) * ldy + n), w, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) red[warp][i][lane] = acc[i];
  __syncthreads();
  for (int e = threadIdx.x; e < 16 * 32; e += 256) {
    const int i = e >> 5, l = e & 31, kk = blockIdx.x * 32 + l;
    if (i >= mrows || kk >= K) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][i][l];
    float* o = dX + (size_t)(m0 + i) * ldx + kk;
    *o = accumulate ? *o + s : s;
  }
}

// g[r, :] += d LayerNorm(x[r, :]) applied to the upstream gradient dy[r, :] (d = 512, one warp per row).
// With `mod` (TimeBlock, cross_attention.py:426-439): the forward is t = SiLU(LN(x) (1 + scale) + shift) and `dy` is
// the gradient with respect to t.
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, const float* __restrict__ mod,
                                                     const float* __restrict__ dy, float* __restrict__ g, int rows) {
  pdl_sync();
  constexpr int D = 512, PER = D / 32;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + (size_t)r * D;
  float xv[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { xv[i] = xr[lane + 32 * i]; s += xv[i]; }
  const float mean = warp_sum(s) * (1.0f / D);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float c = xv[i] - mean; v = fmaf(c, c, v); }
  const float rstd = rsqrtf(warp_sum(v) * (1.0f / D) + 1e-5f);
  float gd[PER], xh[PER];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    xh[i] = (xv[i] - mean) * rstd;
    float up = dy[(size_t)r * D + c];
    if (mod) {
      const float sc = 1.0f + mod[c];
      const float y = fmaf(fmaf(xh[i], gamma[c], beta[c]), sc, mod[D + c]);
      const float sg = 1.0f / (1.0f + expf(-y));
      up = up * (sg * (1.0f + y * (1.0f - sg))) * sc;        // d SiLU(y) / dy, then the modulation scale
    }
    gd[i] = up * gamma[c];
    s1 += gd[i];
    s2 = fmaf(gd[i], xh[i], s2);
  }
  s1 = warp_sum(s1) * (1.0f / D);
  s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    g[(size_t)r * D + c] += rstd * (gd[i] - s1 - xh[i] * s2);
  }
}

// dz = df * GELU'(z), exact-erf GELU (F.gelu default): Phi(z) + z phi(z).
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ z, float* __restrict__ df, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = z[i];
  const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
  df[i] *= cdf + v * pdf;
}

// Self-attention backward for one (batch entry, head): L <= 16 tokens, head_dim 128, no key padding.
// qkv rows [R, 3 d] (q | k | v, head h at columns h * 128), dO rows [R, d]; writes dqkv rows [R, 3 d].
constexpr int SB_L = 16, SB_HD = 128, SB_P = SB_HD + 1;
__global__ void __launch_bounds__(256) mha_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                      float* __restrict__ dqkv, int L, int d) {
  pdl_sync();
  __shared__ float Q[SB_L][SB_P], Kk[SB_L][SB_P], V[SB_L][SB_P], G[SB_L][SB_P];
  __shared__ float P[SB_L][SB_L + 1], dS[SB_L][SB_L + 1];
  const int b = blockIdx.x, h = blockIdx.y, t = threadIdx.x;
  const float scale = rsqrtf((float)SB_HD);
  for (int e = t; e < SB_L * SB_HD; e += 256) {
    const int i = e / SB_HD, c = e % SB_HD;
    float q = 0.f, k = 0.f, v = 0.f, g = 0.f;
    if (i < L) {
      const float* row = qkv + (size_t)(b * L + i) * 3 * d + h * SB_HD + c;
      q = row[0]; k = row[d]; v = row[2 * d];
      g = dO[(size_t)(b * L + i) * d + h * SB_HD + c];
    }
    Q[i][c] = q; Kk[i][c] = k; V[i][c] = v; G[i][c] = g;
  }
  __syncthreads();
  const int i = t >> 4, j = t & 15;
  float sc = 0.f, dp = 0.f;
  for (int c = 0; c < SB_HD; ++c) { sc = fmaf(Q[i][c], Kk[j][c], sc); dp = fmaf(G[i][c], V[j][c], dp); }
  sc = (i < L && j < L) ? sc * scale : -INFINITY;
  // softmax over j inside each group of 16 lanes
  float mx = sc;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float e = (i < L && j < L) ? expf(sc - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float p = (i < L && j < L) ? e / sum : 0.f;
  float dot = p * dp;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  P[i][j] = p;
  dS[i][j] = p * (dp - dot);
  __syncthreads();
  for (int e2 = t; e2 < SB_L * SB_HD; e2 += 256) {
    const int r = e2 / SB_HD, c = e2 % SB_HD;
    if (r >= L) continue;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int m = 0; m < L; ++m) {
      dq = fmaf(dS[r][m], Kk[m][c], dq);
      dk = fmaf(dS[m][r], Q[m][c], dk);
      dv = fmaf(P[m][r], G[m][c], dv);
    }
    float* row = dqkv + (size_t)(b * L + r) * 3 * d + h * SB_HD + c;
    row[0] = dq * scale; row[d] = dk * scale; row[2 * d] = dv;
  }
}

// Backward of one folded single-head cross-attention (batch entry, stream): forward (attention.cu cross_kernel)
//   s_ij = q_i . xhat_j (+ mask),  P = softmax_j(s),  u_i = sum_j P_ij xhat_j
// given du (gradient of the loss w.r.t. u, a 512-column slice of d cat), the probabilities P saved by the forward and
// optionally dAtt (direct gradient on P: the word-excitation loss); writes dq (the same slice of d qx).
// grid (n_batch, 5), 256 threads; shared memory: du [16][512] + dS [16][len].
__global__ void __launch_bounds__(256) cross_bwd_kernel(const float* __restrict__ dcat, const float* __restrict__ mem_hat,
                                                        float* __restrict__ dqx, CrossArgs a, const float* __restrict__ d_att,
                                                        int att_stream, int n_tokens) {
  pdl_sync();
  extern __shared__ float sm[];
  constexpr int D = 512;
  const int bs = blockIdx.x, x = blockIdx.y;
  const int M = a.len[x], ld = CFB_N_STREAMS * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dU = sm;              // [16][512]
  float* dSm = dU + 16 * D;    // [16][M]
  const int slot = a.slot[x] ? a.slot[x][bs] : bs;
  const float* mem = mem_hat + ((size_t)a.row_base[x] + (size_t)slot * M) * D;
  const float* Pm = a.att[x] + (long long)bs * a.att_batch_stride[x];                  // [16][M] of this layer
  const float* dA = (d_att != nullptr && x == att_stream) ? d_att + (long long)bs * a.att_batch_stride[x] : nullptr;
  for (int e = threadIdx.x; e < 16 * D; e += 256) {
    const int i = e / D, c = e % D;
    dU[e] = i < n_tokens ? dcat[(size_t)(bs * n_tokens + i) * ld + x * D + c] : 0.f;
  }
  __syncthreads();
  // dP_ij = du_i . xhat_j (+ dAtt_ij): warp per key, lanes over columns
  for (int j = warp; j < M; j += 8) {
    float kv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) kv[i] = mem[(size_t)j * D + lane + 32 * i];
    for (int qi = 0; qi < n_tokens; ++qi) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc = fmaf(dU[qi * D + lane + 32 * i], kv[i], acc);
      acc = warp_sum(acc);
      if (lane == 0) dSm[qi * M + j] = acc + (dA ? dA[(size_t)qi * M + j] : 0.f);
    }
  }
  __syncthreads();
  // dS = P (dP - sum_j dP P): warp per query row
  for (int qi = warp; qi < n_tokens; qi += 8) {
    float dot = 0.f;
    for (int j = lane; j < M; j += 32) dot = fmaf(dSm[qi * M + j], Pm[(size_t)qi * M + j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < M; j += 32) dSm[qi * M + j] = Pm[(size_t)qi * M + j] * (dSm[qi * M + j] - dot);
  }
  __syncthreads();
  // dq_i = sum_j dS_ij xhat_j: thread t owns columns t and t + 256
  float acc0[16], acc1[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
  for (int j = 0; j < M; ++j) {
    const float k0 = mem[(size_t)j * D + threadIdx.x], k1 = mem[(size_t)j * D + 256 + threadIdx.x];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float s = i < n_tokens ? dSm[i * M + j] : 0.f;
      acc0[i] = fmaf(s, k0, acc0[i]);
      acc1[i] = fmaf(s, k1, acc1[i]);
    }
  }
  for (int i = 0; i < n_tokens; ++i) {
    float* o = dqx + (size_t)(bs * n_tokens + i) * ld + x * D;
    o[threadIdx.x] = acc0[i];
    o[256 + threadIdx.x] = acc1[i];
  }
}

}  // namespace

int linear_bwd(const float* dY, int ldy, const float* W, int ldw, float* dX, int ldx, int M, int N, int K, int accumulate,
               cudaStream_t st) {
  CFB_CHECK(M > 0 && N > 0 && K > 0, "linear_bwd: empty problem %dx%dx%d", M, N, K);
  launch_k(linear_bwd_kernel, dim3(ceil_div(K, 32), ceil_div(M, 16)), dim3(256), 0, st, dY, ldy, W, ldw, dX, ldx, M, N, K,
           accumulate);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int ln_bwd(const float* x, const float* gamma, const float* beta, const float* mod, const float* dy, float* g, int rows, int d,
           cudaStream_t st) {
  CFB_CHECK(d == 512, "ln_bwd: d=%d unsupported (512)", d);
  launch_k(ln_bwd_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, st, x, gamma, beta, mod, dy, g, rows);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int gelu_bwd(const float* z, float* df, long long n, cudaStream_t st) {
  launch_k(gelu_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, z, df, n);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int mha_bwd(const float* qkv, const float* dO, float* dqkv, int n_batch, int L, int n_heads, int d, cudaStream_t st) {
  CFB_CHECK(L <= SB_L && d == n_heads * SB_HD, "mha_bwd: L=%d d=%d heads=%d unsupported", L, d, n_heads);
  launch_k(mha_bwd_kernel, dim3(n_batch, n_heads), dim3(256), 0, st, qkv, dO, dqkv, L, d);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

int cross_bwd(const float* dcat, const float* mem_hat, float* dqx, const CrossArgs& a, const float* d_att, int att_stream,
              int n_batch, int n_tokens, cudaStream_t st) {
  int maxM = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    CFB_CHECK(a.att[x] != nullptr, "cross_bwd: the forward's probabilities of stream %d were not saved", x);
    if (a.len[x] > maxM) maxM = a.len[x];
  }
  const size_t smem = ((size_t)16 * 512 + (size_t)16 * maxM) * sizeof(float);
  CFB_CHECK(smem <= 200 * 1024, "cross_bwd: %d memory tokens exceed the shared-memory budget", maxM);
  if (smem > 48 * 1024)   // a per-device attribute; only long memories need it, setting it again is cheap
    CFB_CUDA(cudaFuncSetAttribute(cross_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  launch_k(cross_bwd_kernel, dim3(n_batch, CFB_N_STREAMS), dim3(256), smem, st, dcat, mem_hat, dqx, a, d_att, att_stream,
           n_tokens);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

}  // namespace cfb
