// ConvoFusionVae.decode (vae.py:268-372): two SkipTransformerDecoders (body / hands, d=128, 5 layers,
// 2 heads, FFN 1024) over 128 frame queries per clip, cross-attending to 8 latent tokens each, then
// the final linears to 69 + 120 joint features and zeroing of padded frames.
// Rows are clip-major (row = clip * n_frames + frame); the residual stream stays fp32.
#include "common.cuh"
#include "kernels.cuh"
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace cfb {
// 16-bit VAE handles: weights (packed as fp16 by the host when this is set) and activations are fp16, not bf16 -- 11
// instead of 8 significant bits at the same bytes and MMA rate; the VAE's values are O(1)-O(100) and every store
// clamps to the fp16 range.  cfb_set_vae_f16 / env CFB_VAE_F16; read when a handle is created (the packer asks
// cfb_get_vae_f16 at the same moment).  Default 1.
int g_vae_f16 = getenv("CFB_VAE_F16") ? atoi(getenv("CFB_VAE_F16")) : 1;
int init_gemm_tc_kernels();
int init_attention_kernels();
}

using namespace cfb;

struct VaeBuf {
  void* p = nullptr; size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return CFB_OK;
    if (p) CFB_CUDA(cudaFree(p));
    p = nullptr; cap = 0;
    CFB_CUDA(cudaMalloc(&p, bytes + 256));
    cap = bytes + 256;
    return CFB_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

struct cfb_vae {
  cfb_vae_weights w;
  std::vector<cfb_vae_layer> layers[2];
  std::vector<cfb_vae_enc_layer> enc_layers[2];
  int d, L, H, ff, prec;
  int f16 = 0;     // 16-bit handle whose weights were packed as fp16 (g_vae_f16 at creation)
  VaeBuf h, a, qkv, q, kv, mem, f, cat, skip[2], lens;
};

namespace {

// T = __half is the fp16 form of the 16-bit handle: same buffers, fp16 payloads, fp16 x fp16 GEMMs / attention
template <typename T> constexpr int is_h() { return std::is_same<T, __half>::value ? 1 : 0; }
template <typename T>
int ln16(const float* x, const float* g, const float* b, T* out, int rows, int d, cudaStream_t st) {
  if constexpr (std::is_same<T, __half>::value)
    return ln_rows<bf16>(x, g, b, nullptr, nullptr, 0, reinterpret_cast<bf16*>(out), rows, d, st, 3);
  else return ln_rows<T>(x, g, b, nullptr, nullptr, 0, out, rows, d, st);
}

template <typename T>
int vae_layer(cfb_vae* v, const cfb_vae_layer& w, int n_clips, int n_frames, int n_chunks, cudaStream_t st) {
  // cross_attention.py:361-382 TransformerDecoderLayer.forward_pre
  const int R = n_clips * n_frames, Rm = n_clips * n_chunks, d = v->d, tb = sizeof(T) == 2;
  float* h = v->h.as<float>();
  T* a = v->a.as<T>();
  auto lin_T = [&](const void* A, int rows, int K, const void* W, const float* b, void* out, int N, int act) {
    Epilogue ep{}; ep.bias = b; ep.bias_period = 1; ep.act = act; ep.out_bf16 = tb; ep.out = out; ep.ldo = N; ep.replicate = 1;
    ep.ab_f16 = ep.out_f16 = is_h<T>();
    return gemm(A, tb, K, W, tb, K, rows, N, K, 0, ep, st);
  };
  auto lin_res = [&](const void* A, int K, const void* W, const float* b) {
    Epilogue ep{}; ep.bias = b; ep.bias_period = 1; ep.accumulate = 1; ep.out = h; ep.ldo = d; ep.replicate = 1;
    ep.ab_f16 = is_h<T>();
    return gemm(A, tb, K, W, tb, K, R, d, K, 0, ep, st);
  };
  T* qkv = v->qkv.as<T>();
  CFB_TRY(ln16<T>(h, w.ln1_g, w.ln1_b, a, R, d, st));
  CFB_TRY(lin_T(a, R, d, w.w_in, w.b_in, qkv, 3 * d, 0));
  CFB_TRY(mha<T>(qkv, 3 * d, qkv + d, qkv + 2 * d, 3 * d, a, d, n_clips, n_frames, n_frames, v->H, d / v->H,
                 v->lens.as<int>(), st));                      // tgt_key_padding_mask = ~mask (vae.py:326)
  CFB_TRY(lin_res(a, d, w.w_so, w.b_so));
  CFB_TRY(ln16<T>(h, w.ln2_g, w.ln2_b, a, R, d, st));
  CFB_TRY(lin_T(a, R, d, w.w_q, w.b_q, v->q.p, d, 0));
  CFB_TRY(lin_T(v->mem.p, Rm, d, w.w_kv, w.b_kv, v->kv.p, 2 * d, 0));
  T* kv = v->kv.as<T>();
  CFB_TRY(mha<T>(v->q.as<T>(), d, kv, kv + d, 2 * d, a, d, n_clips, n_frames, n_chunks, v->H, d / v->H, nullptr, st));
  CFB_TRY(lin_res(a, d, w.w_co, w.b_co));
  CFB_TRY(ln16<T>(h, w.ln3_g, w.ln3_b, a, R, d, st));
  CFB_TRY(lin_T(a, R, d, w.w_ff1, w.b_ff1, v->f.p, v->ff, CFB_ACT_GELU));
  return lin_res(v->f.p, v->ff, w.w_ff2, w.b_ff2);
}

template <typename T>
int vae_decode_part(cfb_vae* v, int part, const float* z_part, int n_clips, int n_chunks, int n_frames, float* out,
                    int out_ld, int col_off, cudaStream_t st) {
  const cfb_vae_decoder& dw = v->w.part[part];
  const int R = n_clips * n_frames, d = v->d, tb = sizeof(T) == 2;
  const int nb = (v->L - 1) / 2;
  float* h = v->h.as<float>();
  CFB_TRY(add_pe<float>(nullptr, v->w.pe_query, h, n_clips, n_frames, d, st));            // vae.py:277,321
  CFB_TRY(add_pe<T>(z_part, v->w.pe_mem, v->mem.as<T>(), n_clips, n_chunks, d, st));      // vae.py:322,331
  int li = 0;
  for (int i = 0; i < nb; ++i) {                                                          // cross_attention.py:99-105
    CFB_TRY(vae_layer<T>(v, v->layers[part][li++], n_clips, n_frames, n_chunks, st));
    CFB_CUDA(cudaMemcpyAsync(v->skip[i].p, h, (size_t)R * d * 4, cudaMemcpyDeviceToDevice, st));
  }
  CFB_TRY(vae_layer<T>(v, v->layers[part][li++], n_clips, n_frames, n_chunks, st));       // middle_block
  for (int i = 0; i < nb; ++i) {                                                          // :113-120
    CFB_TRY(concat2<T>(h, v->skip[nb - 1 - i].as<float>(), v->cat.as<T>(), R, d, st));
    Epilogue ep{}; ep.bias = dw.b_skip[i]; ep.bias_period = 1; ep.out = h; ep.ldo = d; ep.replicate = 1;
    ep.ab_f16 = is_h<T>();
    CFB_TRY(gemm(v->cat.p, tb, 2 * d, dw.w_skip[i], tb, 2 * d, R, d, 2 * d, 0, ep, st));
    CFB_TRY(vae_layer<T>(v, v->layers[part][li++], n_clips, n_frames, n_chunks, st));
  }
  CFB_TRY(ln16<T>(h, dw.lnf_g, dw.lnf_b, v->a.as<T>(), R, d, st));   // :122-123
  Epilogue ep{}; ep.bias = dw.b_final; ep.bias_period = 1; ep.out = out + col_off; ep.ldo = out_ld; ep.replicate = 1;
  ep.ab_f16 = is_h<T>();
  return gemm(v->a.p, tb, d, dw.w_final, tb, d, R, dw.n_out, d, 0, ep, st);               // vae.py:352-353
}

// cross_attention.py:288-300 TransformerEncoderLayer.forward_pre over R = n * L token rows (h fp32 residual).
template <typename T>
int vae_enc_layer(cfb_vae* v, const cfb_vae_enc_layer& w, int n, int L, cudaStream_t st) {
  const int R = n * L, d = v->d, tb = sizeof(T) == 2;
  float* h = v->h.as<float>();
  T* a = v->a.as<T>();
  T* qkv = v->qkv.as<T>();
  auto lin_T = [&](const void* A, int K, const void* W, const float* b, void* out, int N, int act) {
    Epilogue ep{}; ep.bias = b; ep.bias_period = 1; ep.act = act; ep.out_bf16 = tb; ep.out = out; ep.ldo = N; ep.replicate = 1;
    ep.ab_f16 = ep.out_f16 = is_h<T>();
    return gemm(A, tb, K, W, tb, K, R, N, K, 0, ep, st);
  };
  auto lin_res = [&](const void* A, int K, const void* W, const float* b) {
    Epilogue ep{}; ep.bias = b; ep.bias_period = 1; ep.accumulate = 1; ep.out = h; ep.ldo = d; ep.replicate = 1;
    ep.ab_f16 = is_h<T>();
    return gemm(A, tb, K, W, tb, K, R, d, K, 0, ep, st);
  };
  CFB_TRY(ln16<T>(h, w.ln1_g, w.ln1_b, a, R, d, st));
  CFB_TRY(lin_T(a, d, w.w_in, w.b_in, qkv, 3 * d, 0));
  CFB_TRY(mha<T>(qkv, 3 * d, qkv + d, qkv + 2 * d, 3 * d, a, d, n, L, L, v->H, d / v->H, v->lens.as<int>(), st));
  CFB_TRY(lin_res(a, d, w.w_so, w.b_so));
  CFB_TRY(ln16<T>(h, w.ln2_g, w.ln2_b, a, R, d, st));
  CFB_TRY(lin_T(a, d, w.w_ff1, w.b_ff1, v->f.p, v->ff, CFB_ACT_GELU));
  return lin_res(v->f.p, v->ff, w.w_ff2, w.b_ff2);
}

template <typename T>
int vae_encode_part(cfb_vae* v, int part, const float* feats, int n_feat, int n, float* mu, float* sd, cudaStream_t st) {
  const cfb_vae_encoder& ew = v->w.enc[part];
  const int chunk = 16, n_tok = 2, L = n_tok + chunk, R = n * L, d = v->d, tb = sizeof(T) == 2;
  const int nb = (v->L - 1) / 2;
  float* h = v->h.as<float>();
  float* emb = v->skip[1].as<float>();
  {                                                                                      // vae.py:194-198 skel embedding
    Epilogue ep{}; ep.bias = ew.b_emb; ep.bias_period = 1; ep.out = emb; ep.ldo = d; ep.replicate = 1;
    CFB_TRY(gemm(feats + ew.col0, 0, n_feat, ew.w_emb, 0, ew.n_in, n * chunk, d, ew.n_in, 0, ep, st));
  }
  CFB_TRY(enc_assemble(emb, ew.tokens, v->w.pe_enc, h, n, n_tok, chunk, d, st));
  int li = 0;
  for (int i = 0; i < nb; ++i) {                                                         // cross_attention.py:47-51
    CFB_TRY(vae_enc_layer<T>(v, v->enc_layers[part][li++], n, L, st));
    CFB_CUDA(cudaMemcpyAsync(v->skip[i].p, h, (size_t)R * d * 4, cudaMemcpyDeviceToDevice, st));
  }
  CFB_TRY(vae_enc_layer<T>(v, v->enc_layers[part][li++], n, L, st));
  for (int i = 0; i < nb; ++i) {                                                         // :56-60
    CFB_TRY(concat2<T>(h, v->skip[nb - 1 - i].as<float>(), v->cat.as<T>(), R, d, st));
    Epilogue ep{}; ep.bias = ew.b_skip[i]; ep.bias_period = 1; ep.out = h; ep.ldo = d; ep.replicate = 1;
    ep.ab_f16 = is_h<T>();
    CFB_TRY(gemm(v->cat.p, tb, 2 * d, ew.w_skip[i], tb, 2 * d, R, d, 2 * d, 0, ep, st));
    CFB_TRY(vae_enc_layer<T>(v, v->enc_layers[part][li++], n, L, st));
  }
  float* y = v->skip[0].as<float>();
  CFB_TRY(ln_rows<float>(h, ew.lnf_g, ew.lnf_b, nullptr, nullptr, 0, y, R, d, st));      // :62-63
  return enc_dist(y, mu, sd, n, L, d, st);
}

}  // namespace

extern "C" {

int cfb_vae_create(const cfb_vae_weights* w, cfb_vae** out) {
  CFB_CHECK(w && out, "cfb_vae_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: convofusion_b200 has no CPU fallback");
    return CFB_ERR_NO_DEVICE;
  }
  CFB_CHECK(w->d_model == 128, "vae d_model %d unsupported (128)", w->d_model);
  CFB_CHECK(w->n_layers % 2 == 1 && w->n_layers >= 1 && w->n_layers <= 9, "vae n_layers must be odd and <= 9");
  CFB_CHECK(w->precision == CFB_F32 || w->precision == CFB_BF16, "bad precision");
  cfb_vae* v = new cfb_vae();
  v->w = *w;
  for (int p = 0; p < 2; ++p) {
    CFB_CHECK(w->part[p].layers != nullptr, "vae part %d has no layers", p);
    v->layers[p].assign(w->part[p].layers, w->part[p].layers + w->n_layers);
    v->w.part[p].layers = v->layers[p].data();
    if (w->enc[p].layers) {
      v->enc_layers[p].assign(w->enc[p].layers, w->enc[p].layers + w->n_layers);
      v->w.enc[p].layers = v->enc_layers[p].data();
    }
  }
  v->d = w->d_model; v->L = w->n_layers; v->H = w->n_heads; v->ff = w->ff_size; v->prec = w->precision;
  v->f16 = (w->precision == CFB_BF16 && g_vae_f16) ? 1 : 0;
  int rc = init_gemm_tc_kernels();
  if (rc == CFB_OK) rc = init_attention_kernels();
  if (rc != CFB_OK) { delete v; return rc; }
  *out = v;
  return CFB_OK;
}

int cfb_set_vae_f16(int enabled) {
  g_vae_f16 = enabled ? 1 : 0;
  return CFB_OK;
}
int cfb_get_vae_f16(void) { return g_vae_f16; }

void cfb_vae_destroy(cfb_vae* v) {
  if (!v) return;
  VaeBuf* bufs[] = {&v->h, &v->a, &v->qkv, &v->q, &v->kv, &v->mem, &v->f, &v->cat, &v->skip[0], &v->skip[1], &v->lens};
  for (VaeBuf* b : bufs) b->release();
  delete v;
}

int cfb_vae_decode(cfb_vae* v, const float* z, int n_clips, int n_chunks, int n_frames, const int32_t* lengths_host,
                   float* out, cfb_stream stream) {
  CFB_CHECK(v && z && out && lengths_host && n_clips > 0 && n_chunks > 0 && n_frames > 0, "cfb_vae_decode: bad argument");
  CFB_CHECK(n_frames <= v->w.pe_len && n_chunks <= v->w.pe_len, "cfb_vae_decode: sequence exceeds the positional table");
  CFB_CHECK((v->L - 1) / 2 <= 2, "cfb_vae_decode: at most 2 skip connections supported");
  for (int b = 0; b < n_clips; ++b)
    CFB_CHECK(lengths_host[b] >= 1 && lengths_host[b] <= n_frames, "cfb_vae_decode: length[%d]=%d outside [1,%d]", b, lengths_host[b], n_frames);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t R = (size_t)n_clips * n_frames, Rm = (size_t)n_clips * n_chunks, d = v->d, es = v->prec == CFB_BF16 ? 2 : 4;
  CFB_TRY(v->h.reserve(R * d * 4));
  CFB_TRY(v->a.reserve(R * d * es));
  CFB_TRY(v->qkv.reserve(R * 3 * d * es));
  CFB_TRY(v->q.reserve(R * d * es));
  CFB_TRY(v->kv.reserve(Rm * 2 * d * es));
  CFB_TRY(v->mem.reserve(Rm * d * es));
  CFB_TRY(v->f.reserve(R * v->ff * es));
  CFB_TRY(v->cat.reserve(R * 2 * d * es));
  CFB_TRY(v->skip[0].reserve(R * d * 4));
  CFB_TRY(v->skip[1].reserve(R * d * 4));
  CFB_TRY(v->lens.reserve((size_t)n_clips * 4));
  CFB_CUDA(cudaMemcpyAsync(v->lens.p, lengths_host, (size_t)n_clips * 4, cudaMemcpyHostToDevice, st));
  CFB_CUDA(cudaStreamSynchronize(st));   // lengths_host may be a caller temporary
  const int n_out = v->w.part[0].n_out + v->w.part[1].n_out;
  int col = 0;
  for (int p = 0; p < 2; ++p) {
    const float* zp = z + (size_t)p * Rm * d;     // torch.chunk(z, 2, dim=0), vae.py:279
    if (v->prec == CFB_BF16 && v->f16) CFB_TRY(vae_decode_part<__half>(v, p, zp, n_clips, n_chunks, n_frames, out, n_out, col, st));
    else if (v->prec == CFB_BF16) CFB_TRY(vae_decode_part<bf16>(v, p, zp, n_clips, n_chunks, n_frames, out, n_out, col, st));
    else CFB_TRY(vae_decode_part<float>(v, p, zp, n_clips, n_chunks, n_frames, out, n_out, col, st));
    col += v->w.part[p].n_out;
  }
  return mask_frames(out, v->lens.as<int>(), n_clips, n_frames, n_out, st);   // vae.py:362
}

int cfb_vae_encode(cfb_vae* v, const float* features, int n_clips, int n_frames, const int32_t* lengths_host,
                   float* mu_out, float* std_out, float* feats_out, cfb_stream stream) {
  CFB_CHECK(v && features && lengths_host && mu_out && std_out && feats_out && n_clips > 0, "cfb_vae_encode: bad argument");
  CFB_CHECK(n_frames >= 16 && n_frames % 16 == 0, "cfb_vae_encode: n_frames=%d must be a positive multiple of 16", n_frames);
  CFB_CHECK(v->w.enc[0].layers && v->w.enc[1].layers && v->w.pe_enc, "cfb_vae_encode: handle was created without encoder weights");
  CFB_CHECK((v->L - 1) / 2 <= 2, "cfb_vae_encode: at most 2 skip connections supported");
  CFB_CHECK(v->w.pe_len >= 18, "cfb_vae_encode: positional table shorter than 18");
  const int chunk = 16, n_tok = 2, Lt = n_tok + chunk, n_chunks = n_frames / chunk, n = n_clips * n_chunks;
  const int n_feat = v->w.enc[0].n_in + v->w.enc[1].n_in;
  CFB_CHECK(v->w.enc[0].col0 == 0 && v->w.enc[1].col0 == v->w.enc[0].n_in, "cfb_vae_encode: body/hands columns must tile the feature row");
  // key lengths per chunk: the 2 distribution tokens are always valid, then the chunk's share of lengths_to_mask
  std::vector<int> kv(n);
  for (int b = 0; b < n_clips; ++b) {
    CFB_CHECK(lengths_host[b] >= 0 && lengths_host[b] <= n_frames, "cfb_vae_encode: length[%d]=%d outside [0,%d]", b, lengths_host[b], n_frames);
    for (int c = 0; c < n_chunks; ++c) {
      int k = lengths_host[b] - c * chunk;
      kv[b * n_chunks + c] = n_tok + (k < 0 ? 0 : k > chunk ? chunk : k);
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t R = (size_t)n * Lt, d = v->d, es = v->prec == CFB_BF16 ? 2 : 4;
  CFB_TRY(v->h.reserve(R * d * 4));
  CFB_TRY(v->a.reserve(R * d * es));
  CFB_TRY(v->qkv.reserve(R * 3 * d * es));
  CFB_TRY(v->f.reserve(R * v->ff * es));
  CFB_TRY(v->cat.reserve(R * 2 * d * es));
  CFB_TRY(v->skip[0].reserve(R * d * 4));
  CFB_TRY(v->skip[1].reserve(R * d * 4));
  CFB_TRY(v->lens.reserve((size_t)n * 4));
  CFB_CUDA(cudaMemcpyAsync(v->lens.p, kv.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CFB_CUDA(cudaStreamSynchronize(st));
  CFB_TRY(chunk_root(features, feats_out, (long long)n_clips * n_frames, n_feat, chunk, st));   // vae.py:176-186
  for (int p = 0; p < 2; ++p) {
    float* mu = mu_out + (size_t)p * n * d;        // torch.cat((b_mu, h_mu), axis=0), vae.py:254-255
    float* sd = std_out + (size_t)p * n * d;
    if (v->prec == CFB_BF16 && v->f16) CFB_TRY(vae_encode_part<__half>(v, p, feats_out, n_feat, n, mu, sd, st));
    else if (v->prec == CFB_BF16) CFB_TRY(vae_encode_part<bf16>(v, p, feats_out, n_feat, n, mu, sd, st));
    else CFB_TRY(vae_encode_part<float>(v, p, feats_out, n_feat, n, mu, sd, st));
  }
  return CFB_OK;
}

}  // extern "C"
