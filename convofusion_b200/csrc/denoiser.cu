// Denoiser handle: one evaluation (cfb_denoiser_forward) and the whole guided sampling loop
// (cfb_sample) as a sequence of the kernels in gemm_*.cu / rowops.cu / attention.cu / sched.cu.
//
// Data layout in HBM (R = n_batch * n_tokens query rows, batch-major like the reference's
// batch-first [BG,16,128] sample; d = 512):
//   h        float [R, d]      residual stream (always fp32)
//   a        T     [R, d]      current GEMM A operand (LayerNorm / modulation output, attention output)
//   qkv      T     [R, 3d]     self-attention projections
//   qx       T     [R, 5d]     folded cross-attention queries, overwritten in place by P.mem_hat
//   f        T     [R, ff]     GELU(linear1)
//   mem_c    float [Rm, d]     cond + stream embedding + PE for every memory slot (built once per call)
//   mem_hat  T     [Rm, d]     LayerNorm-without-affine of (mem_c + time_emb(t)), rebuilt every step
//   temb / tbmod float [S, d] / [S, L*2*2d]   time embedding and TimeBlock (scale|shift) for all S steps
// T = float (fp32 parity mode) or bf16 (tensor-core mode).
#include "common.cuh"
#include "kernels.cuh"
#include "rowblock.cuh"
#include <vector>
#include <cstring>
#include <cstdlib>

namespace cfb {

int init_gemm_tc_kernels();
int init_attention_kernels();
extern int g_cross_tc;

struct DeviceBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes, unsigned* epoch) {
    if (bytes <= cap) return CFB_OK;
    if (p) CFB_CUDA(cudaFree(p));
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    CFB_CUDA(cudaMalloc(&p, want));
    cap = want;
    if (epoch) ++*epoch;
    return CFB_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

}  // namespace cfb

namespace cfb {
int g_shared_plan = getenv("CFB_PLAN") ? atoi(getenv("CFB_PLAN")) : 1;   // cfb_set_shared_plan / env CFB_PLAN=0
// Row-block kernel (rowblock.cu) for the residual chains of a layer; bit 0: out_proj -> TimeBlock 1 -> norm2,
// bit 1: shared values (+ conditional fuser block) -> TimeBlock 2 -> norm3, bit 2: linear2 -> next norm1.
// cfb_set_rowblock / env CFB_ROWBLOCK.  Default 0 (one kernel per operator): measured on the B200 the row-block
// programs are correct but SLOWER at the reference's batch sizes -- a 128-row block is serialised on one SM (MMA phase,
// then statistics pass, then LayerNorm pass: 63 us per two-GEMM chain, profiles/r02_rowblock_trace.txt) while the
// operator path spreads the same rows over 4 n-tiles x 2 co-resident CTAs.
int g_rowblock = getenv("CFB_ROWBLOCK") ? atoi(getenv("CFB_ROWBLOCK")) : 0;
int g_rb_trace_kind = -1, g_rb_trace_layer = -1;   // debug: which program to trace (cfb_debug_rb_trace_arm)
// fp32 handles: GEMMs of the denoiser as three-way bf16 splits on the tcgen05 tensor cores (gemm_split.cu) instead of
// the CUDA-core FFMA kernel (default on: 1e-4 parity holds on both, the split is 2.4x faster at batch 16).
// cfb_set_fp32_tensor_cores / env CFB_FP32_TC=0 selects the CUDA cores.
int g_fp32_tc = getenv("CFB_FP32_TC") ? atoi(getenv("CFB_FP32_TC")) : 1;
// bf16 handles: which LayerNorm outputs are kept as TWO bf16 terms per value (hi + lo, 16 mantissa bits) for the GEMM
// they feed (two accumulating MMAs per K step against the bf16 weights); bit mask of consumer sites: 1 qkv, 2 the two
// TimeBlock linears, 8 linear1, 16 latent_proj.  Activation rounding is what the guidance combine amplifies (DESIGN.md
// section 2, tools/split_sites.py): latent_proj alone (one 128-column GEMM per evaluation, free) removes 43 % of the
// bf16 mode's deviation from fp32, latent_proj + TimeBlock linears 62 %, every site 68 %.
// cfb_set_bf16_activation_sites / env CFB_BF16_ACT_SITES; default 16: the fp16 form below covers every site for free,
// and latent_proj -- whose output IS eps -- is worth its second term on top of it (0.028 -> 0.017 at -0.8 % throughput).
bool gemm_tc_two_term_ok();   // gemm_tc.cu
// bf16 handles: the LayerNorm outputs feeding qkv, the TimeBlock linears, linear1 and latent_proj are stored as fp16
// instead of bf16 -- 11 instead of 8 significant bits for values that are O(1) by construction (clamped to the fp16
// range anyway) -- and those GEMMs run tcgen05.mma kind::f16 on fp16 operands: the handle keeps fp16 copies of those
// bf16 weights (an exact conversion for every weight above the fp16 subnormal range, so the weight rounding stays the
// bf16 one, which is common to all guidance branches and cancels).  Same bytes, same MMA rate: the activation rounding
// that the guidance weights amplify shrinks 8x at these sites at no cost.  (kind::f16 cannot mix A = f16 with B = bf16:
// illegal instruction.)  cfb_set_bf16_activation_f16 / env CFB_BF16_ACT_F16; default 1.  Sites in g_bf16_act_sites
// keep their two-term bf16 form.
// Bit mask of operand groups: 1 the LayerNorm outputs above, 2 the shared-slot probabilities with their per-step values
// and the per-pair attention output feeding the fuser, 4 norm2's output with the per-step keys (scores / conditional
// queries), 8 q / k / v of the self-attention, 16 queries and memory of the per-pair attention.
int g_bf16_act_f16 = getenv("CFB_BF16_ACT_F16") ? atoi(getenv("CFB_BF16_ACT_F16")) : 31;
// Consumer sites taken back out of the fp16 form (they keep bf16 operands and the bf16 matrices): 1 qkv, 2 TimeBlock
// linears, 4 scores / conditional queries, 8 linear1, 16 latent_proj, 32 out_proj, 64 linear2, 128 fuser.  Env
// CFB_BF16_ACT_F16_EXCLUDE (power / accuracy experiments, profiles/r02_act_sites.txt).
int g_bf16_act_f16_exclude = getenv("CFB_BF16_ACT_F16_EXCLUDE") ? atoi(getenv("CFB_BF16_ACT_F16_EXCLUDE")) : 0;
int g_bf16_act_sites = getenv("CFB_BF16_ACT_SITES") ? atoi(getenv("CFB_BF16_ACT_SITES")) : 16;
}

using namespace cfb;

struct cfb_denoiser {
  cfb_denoiser_weights w;
  std::vector<cfb_denoiser_layer> layers;
  int d, lat, ntok, L, H, ff, prec;
  unsigned epoch = 0;
  DeviceBuf h, a, a2, qkv, qx, f, xin, eps, mem_c, mem_hat, mem_hat_t, tsteps, tsin, t1, temb, tbmod, coef, step, x, inp_noise,
      preseq, slots, masks, uc, sS, sP, zall, z0all, ytall;
  // cached CUDA graph of one sampling step
  cudaGraphExec_t graph_exec = nullptr;
  struct GraphKey {
    unsigned epoch; int n_clips, n_branch, full_last, n_steps, kind, clip, preseq_len; float scale;
    int n_slots[CFB_N_STREAMS], len[CFB_N_STREAMS]; bool has_mask[CFB_N_STREAMS];
    const void *noise, *record, *att[CFB_N_STREAMS];
    int plan[2 + 3 * TC_MAX_GROUPS];
  } graph_key;
  bool graph_valid = false;
  size_t graph_nodes = 0;   // kernel nodes in the captured step (for the launch counter)
  // The legacy default stream cannot be captured: graph mode on it runs on this private stream, fenced by events.
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // Concurrent chains of the sampling step (independent groups of batch entries on forked streams)
  static constexpr int MAX_CHAINS = 8;
  cudaStream_t chain_st[MAX_CHAINS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_CHAINS] = {};
  int n_chains = 1;
  // fp32-accurate tensor-core GEMMs (gemm_split.cu): per-chain scratch for split activations + cache of split weights
  SplitCache* split_cache = nullptr;
  DeviceBuf split_ws;
  size_t split_row_bytes = 0, split_a_bytes = 0, split_w_bytes = 0;
  bool fp32_tc = false;
  int act_f16 = 0;        // bf16: operand groups kept as fp16 (g_bf16_act_f16 at the last reserve_rows)
  struct W16 { const bf16 *w_in, *w_tb1, *w_tb2, *w_ff1, *w_fu, *w_qx, *w_so, *w_ff2; };   // fp16 copies (bf16-typed pointers: 16-bit payloads)
  std::vector<W16> l16;
  const bf16* w_out16 = nullptr;
  const bf16* w_embed16 = nullptr;
  const bf16 *w_zx16[CFB_N_STREAMS] = {}, *w_yx16[CFB_N_STREAMS] = {};   // memory-side pre-projections [L d, d]
  DeviceBuf w16;
  int act_sites = 0;      // bf16: consumer sites whose LayerNorm input is kept as [hi | lo] in `a2` (g_bf16_act_sites)
  int split_scheme = 0;   // g_fp32_tc - 1 at the last reserve_split (1: fp32-accurate; 2, 3: precision-study schemes)
  // row-block programs (rowblock.cu): 3 per layer, built once per (workspace epoch, batch layout)
  DeviceBuf rb_prog, rb_blk;
  std::vector<RbStage> rb_host;
  std::vector<int> rb_blk_host;
  int rb_off[3] = {0, 0, 0}, rb_len[3] = {0, 0, 0};   // first stage / stage count of program kind k of layer 0
  int rb_per_layer = 0, rb_mask = 0;
  struct RbKey { unsigned epoch; int rows, k_tot, mask; } rb_key = {~0u, 0, 0, 0};
  // device-resident schedule tables of the last cfb_sample call (see cfb_sample)
  std::vector<float> sched_ts, sched_coef;
  unsigned sched_epoch = ~0u;
  cudaEvent_t ev_sched = nullptr;
  int chains_override = 0;   // cfb_denoiser_set_chains (0 = CFB_CHAINS or 6)
  // per chain: side stream for the conditional-pair sub-chain of every layer; per step: two streams that run the
  // memory-side pre-projection (keys / values) while the chains are still in their self-attention blocks
  cudaStream_t chain_st2[MAX_CHAINS] = {}, pre_st[2] = {};
  cudaEvent_t ev_a[MAX_CHAINS] = {}, ev_b[MAX_CHAINS] = {}, ev_mh = nullptr, ev_pre[2] = {};
  // word-excitation guidance: activations saved by cfb_denoiser_weg_forward for cfb_denoiser_weg_backward
  DeviceBuf wg_h, wg_qkv, wg_z, wg_p, wg_g, wg_t1, wg_t2;
  struct WegState { bool valid = false; int n_batch = 0, att_stream = 0; cfb::CrossArgs ca; long long p_off[CFB_N_STREAMS]; } weg;
};

namespace {

struct MemLayout {
  int row_base[CFB_N_STREAMS];
  int total_rows;
  int slot_off[CFB_N_STREAMS];   // element offsets into h->slots
  int mask_off[CFB_N_STREAMS];   // byte offsets into h->masks
};

// Shared-slot plan (bf16 only).  When slot tables are given, slot 0 of every stream is the memory that most of the
// guidance batch attends to (the unconditional constant).  For those (row, stream) pairs the algebra is pushed onto
// the memory side once per step:
//   Z_{x,l} = xhat_0x A_{x,l}        (keys in query space)        -> scores  S = norm2(tgt) . [Z_0;..;Z_4]^T + z0
//   Y_{x,l} = xhat_0x G_{x,l}^T      (values in residual space)   -> update  h += softmax(S) . [Y_0;..;Y_4]
// so the 512->2560 query projection and the 2560->512 fuser shrink to two GEMMs with N = 320 and K = 448.
// Pairs on any other slot (one stream per single-modality branch) run in `groups`: contiguous row blocks with
// their own stream weights (grouped tcgen05 GEMMs) around the per-pair attention kernel.
struct SharedPlan {
  bool on = false;
  int s_off[CFB_N_STREAMS], p_off[CFB_N_STREAMS], kp[CFB_N_STREAMS];
  int n_tot = 0, k_tot = 0;
  int n_groups = 0;
  int g_stream[TC_MAX_GROUPS], g_row_start[TC_MAX_GROUPS], g_rows[TC_MAX_GROUPS];
  // Groups whose row blocks overlap (the full-cond branch carries all five streams) accumulate into the same rows
  // of h, so they are issued in separate rounds: within a round all row blocks are disjoint.
  int g_round[TC_MAX_GROUPS], n_rounds = 0;
};

// ---- fp32 on the tensor cores ---------------------------------------------------------------------------------------
// Scratch layout of h->split_ws: [main A arena | side A arena | 2 x MAX_CHAINS W slots | 2 x (A, W) slots of the
// memory-side pre-projection streams].  A arenas are indexed by ABSOLUTE query row (chains own disjoint row ranges).
SplitCtx split_ctx(const cfb_denoiser* h, int row, int chain, bool side) {
  SplitCtx c{};
  if (!h->fp32_tc) return c;
  uint8_t* base = h->split_ws.as<uint8_t>();
  c.a_ws = base + (side ? h->split_a_bytes : 0) + (size_t)row * h->split_row_bytes;
  c.a_ws_bytes = h->split_a_bytes - (size_t)row * h->split_row_bytes;
  c.w_ws = base + 2 * h->split_a_bytes + (size_t)(2 * chain + (side ? 1 : 0)) * h->split_w_bytes;
  c.w_ws_bytes = h->split_w_bytes;
  c.cache = h->split_cache;
  c.scheme = h->split_scheme;
  return c;
}
SplitCtx split_ctx_pre(const cfb_denoiser* h, int which) {
  SplitCtx c{};
  if (!h->fp32_tc) return c;
  uint8_t* base = h->split_ws.as<uint8_t>() + 2 * h->split_a_bytes + (size_t)2 * cfb_denoiser::MAX_CHAINS * h->split_w_bytes;
  c.a_ws = base + (size_t)(2 * which) * h->split_w_bytes; c.a_ws_bytes = h->split_w_bytes;
  c.w_ws = base + (size_t)(2 * which + 1) * h->split_w_bytes; c.w_ws_bytes = h->split_w_bytes;
  c.cache = h->split_cache;
  c.scheme = h->split_scheme;
  return c;
}

int reserve_split(cfb_denoiser* h, int n_batch, const cfb_memory* mem, bool plan) {
  h->fp32_tc = h->prec == CFB_F32 && g_fp32_tc != 0 && g_gemm_backend != CFB_GEMM_SIMT;
  if (!h->fp32_tc) return CFB_OK;
  h->split_scheme = g_fp32_tc >= 1 && g_fp32_tc <= 4 ? g_fp32_tc - 1 : 0;
  if (!h->split_cache) h->split_cache = split_cache_create();
  const size_t R = (size_t)n_batch * h->ntok, d = h->d;
  size_t kmax = h->ff > (int)d ? h->ff : d, rows_w = d, k_w = d;
  int n_tot = 0, k_tot = 0, maxlen = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    n_tot += (mem->len[x] + 31) & ~31; k_tot += (mem->len[x] + 63) & ~63;
    if (mem->len[x] > maxlen) maxlen = mem->len[x];
  }
  if (!plan && kmax < CFB_N_STREAMS * d) kmax = CFB_N_STREAMS * d;     // general path: fuser over all five streams
  if ((size_t)k_tot > kmax) kmax = k_tot;
  if ((size_t)n_tot > rows_w) rows_w = n_tot;
  if ((size_t)((maxlen + 63) & ~63) > rows_w) rows_w = (maxlen + 63) & ~63;
  if ((size_t)k_tot > k_w) k_w = k_tot;
  h->split_row_bytes = 6 * kmax * 2;
  h->split_a_bytes = R * h->split_row_bytes;
  h->split_w_bytes = (rows_w * 6 * k_w * 2 + 255) & ~(size_t)255;
  return h->split_ws.reserve(2 * h->split_a_bytes + (size_t)(2 * cfb_denoiser::MAX_CHAINS + 4) * h->split_w_bytes, &h->epoch);
}


// Embedding of the batch entries [b0, b0 + nb) of the guidance batch (entry e = branch * n_clips + clip reads the
// latents of `clip`): every chain embeds its own rows on its own stream instead of one replicated launch up front.
// h->xin must hold the cast latents (embed_cast).
// 16-bit handles: the input latents and latent_embd as fp16 (group 1 of g_bf16_act_f16)
inline int embed_f16(const cfb_denoiser* h) {
  return (h->prec == CFB_BF16 && (h->act_f16 & 1) && h->w_embed16 != nullptr) ? 1 : 0;
}
template <typename T>
int cast_latents(cfb_denoiser* h, const float* latents, long long n, cudaStream_t st) {
  if (embed_f16(h)) return cast_rows<__half>(latents, h->xin.as<__half>(), n, st);
  return cast_rows<T>(latents, h->xin.as<T>(), n, st);
}
template <typename T>
int embed_cast(cfb_denoiser* h, const float* latents, int n_clips, cudaStream_t st) {
  return cast_latents<T>(h, latents, (long long)n_clips * h->ntok * h->lat, st);
}
template <typename T>
int embed_rows(cfb_denoiser* h, int b0, int nb, int n_clips, cudaStream_t st) {
  const int tb = sizeof(T) == 2;
  for (int e = b0; e < b0 + nb;) {
    const int clip = e % n_clips;
    const int run = (n_clips - clip < b0 + nb - e) ? n_clips - clip : b0 + nb - e;   // stay inside one branch
    Epilogue ep{};
    ep.bias = h->w.tok_bias; ep.bias_period = h->ntok; ep.out = h->h.as<float>() + (size_t)e * h->ntok * h->d;
    ep.ldo = h->d; ep.replicate = 1; ep.ab_f16 = embed_f16(h);
    CFB_TRY(gemm(h->xin.as<T>() + (size_t)clip * h->ntok * h->lat, tb, h->lat,
                 ep.ab_f16 ? (const void*)h->w_embed16 : h->w.w_embed, tb, h->lat, run * h->ntok, h->d, h->lat, 0, ep, st));
    e += run;
  }
  return CFB_OK;
}

template <typename T>
int embed(cfb_denoiser* h, const float* latents, int n_in, int replicate, cudaStream_t st) {
  // denoiser.py:183-187,316-326: latent_embd + body/hand embedding + SineBH positional encoding
  const int rows = n_in * h->ntok;
  CFB_TRY(cast_latents<T>(h, latents, (long long)rows * h->lat, st));
  Epilogue ep{};
  ep.bias = h->w.tok_bias; ep.bias_period = h->ntok; ep.out = h->h.p; ep.ldo = h->d;
  ep.replicate = replicate; ep.rep_stride = (long long)rows * h->d;
  ep.ab_f16 = embed_f16(h);
  return gemm(h->xin.p, sizeof(T) == 2, h->lat, ep.ab_f16 ? (const void*)h->w_embed16 : h->w.w_embed, sizeof(T) == 2, h->lat,
              rows, h->d, h->lat, 0, ep, st);
}

// bf16 handles: queries and memory of the per-pair attention as fp16 (group 16 of g_bf16_act_f16; mma.sync kernel only)
inline int pair_f16(const cfb_denoiser* h) {
  return (h->prec == CFB_BF16 && !h->l16.empty() && (h->act_f16 & 16) && !g_cross_tc) ? 1 : 0;
}

// Per-step memory-side precompute of the shared-slot plan for all layers: keys Z + key bias z0 (which = 0) or values
// Y^T (which = 1).  The two halves are independent and run on separate streams.  T = bf16: tcgen05 GEMMs; T = float
// (parity mode): the same algebra on the CUDA-core GEMM, so the plan itself is checked against the oracle at 1e-4.
template <typename T>
int shared_precompute(cfb_denoiser* h, const SharedPlan& sp, const MemLayout& ml, const int len[CFB_N_STREAMS],
                      int which, cudaStream_t st) {
  const int d = h->d, Ld = h->L * h->d;
  constexpr int tb = sizeof(T) == 2;
  const T* mh = h->mem_hat.as<T>();
  const SplitCtx sc_pre = split_ctx_pre(h, which);
  const SplitCtx* scp = h->fp32_tc ? &sc_pre : nullptr;
  (void)scp;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    const T* m0 = mh + (size_t)ml.row_base[x] * d;          // slot 0 of stream x: [len[x], d]
    if (which == 0) {
      Epilogue ez{}; ez.bias_period = 1; ez.out_bf16 = tb; ez.out = h->zall.as<T>() + (size_t)sp.s_off[x] * Ld; ez.ldo = Ld; ez.replicate = 1;
      ez.out_f16 = tb && (h->act_f16 & 4) && !(g_bf16_act_f16_exclude & 4) && !h->l16.empty();   // the scores GEMM runs on the fp16 norm2 output (run_layers)
      ez.ab_f16 = pair_f16(h);                                   // the memory is fp16 then (mem_hat)
      if constexpr (tb) CFB_TRY(gemm_tc(m0, d, ez.ab_f16 ? h->w_zx16[x] : (const bf16*)h->w.w_zx[x], d, len[x], Ld, d, ez, st));
      else { ez.split = scp; ez.w_static = 1; CFB_TRY(gemm(m0, 0, d, h->w.w_zx[x], 0, d, len[x], Ld, d, 0, ez, st)); }
    } else {
      const int rows_avail = ml.total_rows - ml.row_base[x];
      Epilogue ey{}; ey.bias_period = 1; ey.out_bf16 = tb; ey.out = h->ytall.as<T>() + sp.p_off[x]; ey.ldo = sp.k_tot; ey.replicate = 1;
      ey.out_f16 = tb && (h->act_f16 & 2) && !h->l16.empty();   // the values GEMM runs on fp16 probabilities (run_layers)
      const int w_rows = sp.kp[x] < rows_avail ? sp.kp[x] : rows_avail;   // columns past len[x] meet P == 0
      ey.ab_f16 = pair_f16(h);
      if constexpr (tb) CFB_TRY(gemm_tc(ey.ab_f16 ? h->w_yx16[x] : (const bf16*)h->w.w_yx[x], d, m0, d, Ld, sp.kp[x], d, ey, st, w_rows));
      else {   // columns past w_rows stay zero
        ey.split = scp; ey.a_static = 1;
        CFB_TRY(gemm(h->w.w_yx[x], 0, d, m0, 0, d, Ld, w_rows, d, 0, ey, st));
      }
    }
  }
  if (which == 0)
    return shared_key_bias<T>(mh, h->z0all.as<float>(), h->w.a_zx, ml.row_base, len, sp.s_off, h->L, sp.n_tot, st, pair_f16(h));
  return CFB_OK;
}

// ---- row-block programs -------------------------------------------------------------------------------------------
// Per layer three programs over the same 128-row blocks (stage kinds in rowblock.cuh):
//   kind 0  [load h] out_proj (A = self-attention output) -> LN/modulate/SiLU -> TimeBlock 1 -> norm2      -> h, a
//   kind 1  [load h] conditional fuser block (A = per-pair attention output; blocks with a conditional stream only)
//           -> shared values (A = P, K = k_tot) -> LN/modulate/SiLU -> TimeBlock 2 -> norm3                -> h, a
//   kind 2  [load h] linear2 in K halves (A = GELU(linear1)) -> next layer's norm1 / decoder.norm          -> h, a
// The tensor maps cover whole workspace buffers (absolute rows), so one set serves every chain of the step.
int build_rowblock_programs(cfb_denoiser* h, int n_batch, int n_clips, const SharedPlan& sp, bool want_att, cudaStream_t st) {
  h->rb_mask = 0;
  const int R = n_batch * h->ntok;
  if (h->prec != CFB_BF16 || g_gemm_backend == CFB_GEMM_SIMT || g_rowblock == 0 || (h->act_sites & 11) != 0 || R % 128 != 0 || h->d != 512 ||
      h->ff % 512 != 0 || h->ntok != 16)
    return CFB_OK;
  int mask = g_rowblock & 5;
  const int n_blocks = R / 128;
  std::vector<int> blk(n_blocks, -1);
  bool r2_ok = sp.on && !want_att && n_clips % 8 == 0 && sp.k_tot <= 512 && sp.k_tot % 64 == 0;
  if (r2_ok) {
    for (int z = 0; z < sp.n_groups && r2_ok; ++z) {
      if (sp.g_row_start[z] % 128 != 0 || sp.g_rows[z] % 128 != 0) { r2_ok = false; break; }
      for (int b = sp.g_row_start[z] / 128; b < (sp.g_row_start[z] + sp.g_rows[z]) / 128; ++b) {
        if (blk[b] >= 0) { r2_ok = false; break; }            // two conditional streams on one block
        blk[b] = sp.g_stream[z];
      }
    }
  }
  if (r2_ok) mask |= g_rowblock & 2;
  if (mask == 0) return CFB_OK;
  CFB_TRY(init_rowblock_kernels());
  const cfb_denoiser::RbKey key{h->epoch, R, sp.on ? sp.k_tot : 0, mask};
  if (memcmp(&key, &h->rb_key, sizeof(key)) == 0 && blk == h->rb_blk_host) { h->rb_mask = mask; return CFB_OK; }
  const int d = h->d, nff = h->ff / 512;
  std::vector<RbStage>& pr = h->rb_host;
  pr.clear();
  auto stage = [&](int kind) { RbStage s; memset(&s, 0, sizeof(s)); s.kind = kind; return s; };
  auto gemm_stage = [&](const void* W, int w_rows, int w_cols, int ldw, int w_col0, int K, RbStage* out) -> int {
    RbStage s = stage(RB_GEMM);
    s.K = K; s.N = d; s.w_col0 = w_col0;
    CFB_TRY(rowblock_operand_map(W, w_rows, w_cols, ldw, &s.map_w));
    *out = s;
    return CFB_OK;
  };
  auto a_tma = [&](RbStage& s, const void* A, int cols, int col0) -> int {
    s.a_src = RB_A_TMA; s.a_col0 = col0;
    return rowblock_operand_map(A, R, cols, cols, &s.map_a);
  };
  auto ln_epi = [&](RbStage& s, const float* bias, const float* g, const float* b, const float* mod, bool last) {
    s.epi = RB_EPI_LN; s.bias = bias; s.ln_g = g; s.ln_b = b; s.mod = mod; s.ln_silu = mod != nullptr;
    s.mod_stride = (long long)h->L * 2 * 2 * d;
    if (last) { s.spill = 1; s.store_a = 1; }
  };
  for (int l = 0; l < h->L; ++l) {
    const cfb_denoiser_layer& w = h->layers[l];
    const float* mod1 = h->tbmod.as<float>() + (size_t)(2 * l) * 2 * d;
    const float* mod2 = mod1 + 2 * d;
    const size_t first = pr.size();
    RbStage s;
    // kind 0
    if (l == 0) h->rb_off[0] = (int)pr.size();
    pr.push_back(stage(RB_HLOAD));
    CFB_TRY(gemm_stage(w.w_so, d, d, d, 0, d, &s)); CFB_TRY(a_tma(s, h->a.p, d, 0)); ln_epi(s, w.b_so, w.tb1_g, w.tb1_b, mod1, false); pr.push_back(s);
    CFB_TRY(gemm_stage(w.w_tb1, d, d, d, 0, d, &s)); ln_epi(s, w.b_tb1, w.ln2_g, w.ln2_b, nullptr, true); pr.push_back(s);
    if (l == 0) h->rb_len[0] = (int)pr.size() - h->rb_off[0];
    // kind 1
    if (l == 0) h->rb_off[1] = (int)pr.size();
    if (mask & 2) {
      pr.push_back(stage(RB_HLOAD));
      CFB_TRY(gemm_stage(w.w_fu, d, CFB_N_STREAMS * d, CFB_N_STREAMS * d, 0, d, &s));
      CFB_TRY(a_tma(s, h->uc.p, CFB_N_STREAMS * d, 0)); s.per_stream = 1; pr.push_back(s);
      CFB_TRY(gemm_stage(h->ytall.as<bf16>() + (size_t)l * d * sp.k_tot, d, sp.k_tot, sp.k_tot, 0, sp.k_tot, &s));
      CFB_TRY(a_tma(s, h->sP.p, sp.k_tot, 0)); ln_epi(s, w.b_fu, w.tb2_g, w.tb2_b, mod2, false); pr.push_back(s);
      CFB_TRY(gemm_stage(w.w_tb2, d, d, d, 0, d, &s)); ln_epi(s, w.b_tb2, w.ln3_g, w.ln3_b, nullptr, true); pr.push_back(s);
    }
    if (l == 0) h->rb_len[1] = (int)pr.size() - h->rb_off[1];
    // kind 2
    if (l == 0) h->rb_off[2] = (int)pr.size();
    pr.push_back(stage(RB_HLOAD));
    const bool last_layer = l + 1 == h->L;
    for (int c = 0; c < nff; ++c) {
      CFB_TRY(gemm_stage(w.w_ff2, d, h->ff, h->ff, c * 512, 512, &s));
      CFB_TRY(a_tma(s, h->f.p, h->ff, c * 512));
      if (c + 1 == nff)
        ln_epi(s, w.b_ff2, last_layer ? h->w.lnf_g : h->layers[l + 1].ln1_g, last_layer ? h->w.lnf_b : h->layers[l + 1].ln1_b,
               nullptr, true);
      pr.push_back(s);
    }
    if (l == 0) { h->rb_len[2] = (int)pr.size() - h->rb_off[2]; h->rb_per_layer = (int)(pr.size() - first); }
  }
  CFB_TRY(h->rb_prog.reserve(pr.size() * sizeof(RbStage), &h->epoch));   // captured graphs hold these pointers
  CFB_TRY(h->rb_blk.reserve((size_t)n_blocks * 4, &h->epoch));
  h->rb_blk_host = blk;
  CFB_CUDA(cudaMemcpyAsync(h->rb_prog.p, pr.data(), pr.size() * sizeof(RbStage), cudaMemcpyHostToDevice, st));
  CFB_CUDA(cudaMemcpyAsync(h->rb_blk.p, h->rb_blk_host.data(), (size_t)n_blocks * 4, cudaMemcpyHostToDevice, st));
  CFB_CUDA(cudaStreamSynchronize(st));
  h->rb_key = key;
  h->rb_key.epoch = h->epoch;
  h->rb_mask = mask;
  return CFB_OK;
}

int rowblock_run(cfb_denoiser* h, int layer, int kind, int row0, int R, int rows_total, const int* step_ptr, cudaStream_t st) {
  RbLaunch L{};
  L.prog = h->rb_prog.as<RbStage>() + (size_t)layer * h->rb_per_layer + h->rb_off[kind];
  L.n_stages = h->rb_len[kind];
  L.block0 = row0 / 128; L.n_blocks = R / 128;
  L.blk_stream = h->rb_blk.as<int>();
  L.step_ptr = step_ptr;
  L.h = h->h.as<float>(); L.a = h->a.as<bf16>(); L.rows_total = rows_total;
  L.trace = (g_rb_trace_kind == kind && g_rb_trace_layer == layer);
  return rowblock_launch(L, st);
}

// Optional concurrency inside one chain (all null = strictly sequential on `st`).
struct ChainAux {
  cudaStream_t st2 = nullptr;          // conditional-pair projection + attention run here, next to the shared-slot path
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  cudaEvent_t ev_pre[2] = {nullptr, nullptr};   // memory-side pre-projection finished (waited before layer 0's attention)
};

// Runs the 9 layers + final projection for the batch entries [b0, b0 + n_batch) (a "chain").  Rows of different batch
// entries never interact, so disjoint chains may run concurrently on different streams; they share only read-only
// data (weights, mem_hat, zall / ytall / z0all).  n_batch_total = entries of the whole call (tensor-map extents).
template <typename T>
int run_layers(cfb_denoiser* h, int n_batch, CrossArgs ca, float* const att_base[CFB_N_STREAMS], const int* step_ptr,
               float* eps_out_all, cudaStream_t st, const SharedPlan* sp = nullptr, int b0 = 0, int n_batch_total = 0,
               const ChainAux* aux = nullptr, bool use_rb = false, int chain = 0) {
  if (n_batch_total <= 0) n_batch_total = n_batch;
  const ChainAux no_aux;
  if (!aux) aux = &no_aux;
  const int R = n_batch * h->ntok, d = h->d, row0 = b0 * h->ntok, R_total = n_batch_total * h->ntok;
  const int tb = sizeof(T) == 2;
  // LayerNorm outputs whose consumer site is in `sites` go to `a2` as [hi(64) | lo(64)] per 64 columns (row stride 2 d)
  // and that GEMM reads both terms; everything else uses the dense `a`.  Sites: 1 qkv, 2 TimeBlock linears, 8 linear1,
  // 16 latent_proj (scores / conditional queries address `a` by absolute row and stay plain: no measurable effect).
  const int sites = tb ? h->act_sites : 0;
  // fp16 instead of bf16 operands (g_bf16_act_f16): consumers are tcgen05 GEMMs on fp16 weight copies (Epilogue::ab_f16)
  // or the f16 mma.sync attention kernels.  F16_SITES: LayerNorm consumer sites whose `a` is fp16; 128 = the per-pair
  // attention output feeding the fuser (mma.sync kernel only).
  const int fm = (tb && !h->l16.empty()) ? h->act_f16 : 0;
  const int uc_f16 = ((fm & 2) && !g_cross_tc && !(g_bf16_act_f16_exclude & 128)) ? 1 : 0;
  const int pv_f16 = (fm & 2) ? 1 : 0;                                   // shared-slot probabilities x per-step values
  const int self_f16 = ((fm & 8) && mha_f16_supported(h->ntok, d / h->H)) ? 1 : 0;   // q / k / v of the self-attention
  // 32 = the self-attention output feeding out_proj (fp16 when its kernel runs in fp16), 64 = the GELU output feeding
  // linear2: nothing measurable as activations, but their GEMMs then meet the fp16 weights too
  const int F16_SITES = (((fm & 1) ? (1 | 2 | 8 | 16 | 64) : 0) | ((fm & 4) ? 4 : 0) | (uc_f16 ? 128 : 0) | (self_f16 ? 32 : 0)) &
                        ~g_bf16_act_f16_exclude;
  const int pr_f16 = tb ? pair_f16(h) : 0;                                 // queries / memory of the per-pair attention
  ca.out_f16 = uc_f16; ca.in_f16 = pr_f16;
  const long long mod_stride = (long long)h->L * 2 * 2 * d;
  float* hres = h->h.as<float>() + (size_t)row0 * d;
  T* a_abs = h->a.as<T>();
  T* qx_abs = h->qx.as<T>();
  T* a = a_abs + (size_t)row0 * d;
  T* a2 = sites ? h->a2.as<T>() + (size_t)row0 * 2 * d : nullptr;
  auto ln_to = [&](int site, const float* ln_g, const float* ln_b, const float* mod) {   // LayerNorm for consumer `site`
    // (two-term sites take fp16 terms and the fp16 weights when their group is on: 22 significant bits)
    return (sites & site) ? ln_rows<T>(hres, ln_g, ln_b, mod, mod ? step_ptr : nullptr, mod_stride, a2, R, d, st,
                                       (site & F16_SITES) ? 4 : 2)
                          : ln_rows<T>(hres, ln_g, ln_b, mod, mod ? step_ptr : nullptr, mod_stride, a, R, d, st,
                                       (site & F16_SITES) ? 3 : 1);
  };
  T* qkv = h->qkv.as<T>() + (size_t)row0 * 3 * d;
  T* qx = qx_abs + (size_t)row0 * CFB_N_STREAMS * d;
  T* f = h->f.as<T>() + (size_t)row0 * h->ff;
  float* eps_out = eps_out_all + (size_t)row0 * h->lat;
  ca.bs_offset = b0;
  // row-block kernel for the residual chains: whole 128-row blocks of a plan-driven bf16 step only
  const int rb = (sizeof(T) == 2 && use_rb && row0 % 128 == 0 && R % 128 == 0) ? h->rb_mask : 0;
  // fp32 handles with tensor cores enabled: every GEMM below runs as a three-way bf16 split (gemm_split.cu); the
  // context names this chain's main-stream scratch and the handle's cache of split weights (the side-stream GEMMs of the
  // conditional groups take their own rows of the arena, split_ctx(h, group row, chain, side))
  const SplitCtx sc_main = split_ctx(h, row0, chain, false);
  const SplitCtx* scm = (!tb && h->fp32_tc) ? &sc_main : nullptr;
  // a_from_ln carries a site bit for the precision study (gemm_split scheme 3 + CFB_SPLIT_SITES): 1 qkv, 2 TimeBlock
  // linears, 4 scores / conditional queries, 8 linear1, 16 latent_proj; operands that are not LayerNorm outputs:
  // 32 out_proj (self-attention output), 64 linear2 (GELU output), 128 fuser (per-pair attention output), 256 shared
  // values (probabilities)
  // W16: the fp16 copy of W (null: none, e.g. fp32 handles)
  auto lin_T = [&](int K, const void* W, const void* W16, const float* b, void* out, int N, int act, int site,
                   int out_f16 = 0) {   // A = LayerNorm(h) for `site`
    Epilogue ep{}; ep.bias = b; ep.bias_period = 1; ep.act = act; ep.out_bf16 = tb; ep.out = out; ep.ldo = N; ep.replicate = 1;
    ep.out_f16 = out_f16;
    ep.split = scm; ep.w_static = 1; ep.a_from_ln = site;
    const int at = (sites & site) ? 2 : 1;
    ep.a_terms = at; ep.ab_f16 = (site & F16_SITES) ? 1 : 0;
    if (ep.ab_f16) W = W16;
    return gemm(at == 2 ? (const void*)a2 : (const void*)a, tb, at * K, W, tb, K, R, N, K, 0, ep, st);
  };
  // Residual update followed by the LayerNorm that feeds the next GEMM (optionally with TimeBlock modulation).  Three
  // ways of running that LayerNorm inside the producing GEMM were measured slower than the separate row kernel
  //  and are gone; the row-block kernel (rowblock.cu) is what fuses them now.
  // a_site: 2 = A is the TimeBlock LayerNorm output, >= 32 = A is not a LayerNorm output (attention / GELU / fuser
  // operand, always dense bf16); next_site: the consumer of the LayerNorm run here
  auto lin_res_ln = [&](const void* A, int K, const void* W, const void* W16, const float* b, const float* ln_g,
                        const float* ln_b, const float* mod, int a_site, int next_site) {
    Epilogue ep{}; ep.bias = b; ep.bias_period = 1; ep.accumulate = 1; ep.out = hres; ep.ldo = d; ep.replicate = 1;
    ep.split = scm; ep.w_static = 1; ep.a_from_ln = a_site;
    const int at = (a_site && (sites & a_site)) ? 2 : 1;
    ep.a_terms = at; ep.ab_f16 = (a_site & F16_SITES) ? 1 : 0;
    if (ep.ab_f16) W = W16;
    CFB_TRY(gemm(at == 2 ? (const void*)a2 : A, tb, at * K, W, tb, K, R, d, K, 0, ep, st));
    CFB_TRY(ln_to(next_site, ln_g, ln_b, mod));
    return (int)CFB_OK;
  };
  CFB_TRY(ln_to(1, h->layers[0].ln1_g, h->layers[0].ln1_b, nullptr));
  for (int l = 0; l < h->L; ++l) {
    const cfb_denoiser_layer& w = h->layers[l];
    const float* mod1 = h->tbmod.as<float>() + (size_t)(2 * l) * 2 * d;
    const float* mod2 = mod1 + 2 * d;
    // self-attention block (cross_attention.py:568-572); a = norm1(h) on entry
    const cfb_denoiser::W16 w16 = fm ? h->l16[l] : cfb_denoiser::W16{};
    CFB_TRY(lin_T(d, w.w_in, w16.w_in, w.b_in, qkv, 3 * d, 0, 1, self_f16));
    if (self_f16) {
      if constexpr (tb)
        CFB_TRY(mha_f16(qkv, 3 * d, qkv + d, qkv + 2 * d, 3 * d, a, d, n_batch, h->ntok, h->ntok, h->H, d / h->H, st,
                        (F16_SITES & 32) ? 1 : 0));
    } else {
      CFB_TRY(mha<T>(qkv, 3 * d, qkv + d, qkv + 2 * d, 3 * d, a, d, n_batch, h->ntok, h->ntok, h->H, d / h->H, nullptr, st));
    }
    if (rb & 1) {   // out_proj -> time_block1 -> norm2 on resident rows (rowblock.cu)
      CFB_TRY(rowblock_run(h, l, 0, row0, R, R_total, step_ptr, st));
    } else {
      CFB_TRY(lin_res_ln(a, d, w.w_so, w16.w_so, w.b_so, w.tb1_g, w.tb1_b, mod1, 32, 2));              // + time_block1 prologue (:575)
      CFB_TRY(lin_res_ln(a, d, w.w_tb1, w16.w_tb1, w.b_tb1, w.ln2_g, w.ln2_b, nullptr, 2, 4));       // + norm2 (:578)
    }
    // five cross-attentions + att_fuser (:578-652), folded; a = norm2(h)
    for (int x = 0; x < CFB_N_STREAMS; ++x)
      ca.att[x] = att_base && att_base[x] ? att_base[x] + (size_t)l * h->ntok * ca.len[x] : nullptr;
    bool shared_done = false, rb2_done = false;
    if (sp && sp->on) {
      const int Ld = h->L * d;
      if (l == 0)
        for (int i = 0; i < 2; ++i)
          if (aux->ev_pre[i]) CFB_CUDA(cudaStreamWaitEvent(st, aux->ev_pre[i], 0));
      // conditional pairs: own-stream projection + per-pair attention (on the side stream when there is one: they
      // only read `a` / mem_hat and write qx / uc), then -- after the shared path -- the own-stream fuser blocks
      struct Grp { int x, lo, rows, round; };
      Grp grp[TC_MAX_GROUPS];
      int ng = 0;
      T* uc = h->uc.as<T>();
      float* h_abs = h->h.as<float>();
      for (int z = 0; z < sp->n_groups; ++z) {   // groups carry ABSOLUTE rows; a chain takes the part inside its range
        const int lo = sp->g_row_start[z] > row0 ? sp->g_row_start[z] : row0;
        const int hi_g = sp->g_row_start[z] + sp->g_rows[z], hi = hi_g < row0 + R ? hi_g : row0 + R;
        if (hi <= lo) continue;
        grp[ng++] = Grp{sp->g_stream[z], lo, hi - lo, sp->g_round[z]};
      }
      const bool side = ng > 0 && aux->st2 != nullptr;
      cudaStream_t sc = side ? aux->st2 : st;
      if (ng > 0) {
        if (side) {
          CFB_CUDA(cudaEventRecord(aux->ev_a, st));
          CFB_CUDA(cudaStreamWaitEvent(sc, aux->ev_a, 0));
        }
        Epilogue eq{}; eq.bias_period = 1; eq.out_bf16 = tb; eq.ldo = CFB_N_STREAMS * d; eq.replicate = 1;
        eq.ab_f16 = (F16_SITES & 4) ? 1 : 0; eq.out_f16 = pr_f16;
        if constexpr (sizeof(T) == 2) {
          const bf16* wqx = (F16_SITES & 4) ? w16.w_qx : (const bf16*)w.w_qx;
          TcGroup gq[TC_MAX_GROUPS];
          for (int z = 0; z < ng; ++z)
            gq[z] = TcGroup{a_abs, wqx + (size_t)grp[z].x * d * d, w.b_qx + grp[z].x * d, qx_abs + grp[z].x * d,
                            grp[z].lo, grp[z].rows};
          CFB_TRY(gemm_tc_grouped(gq, ng, R_total, d, d, d, d, eq, sc));
        } else {
          for (int z = 0; z < ng; ++z) {
            Epilogue e1 = eq; e1.bias = w.b_qx + grp[z].x * d;
            e1.out = qx_abs + (size_t)grp[z].lo * CFB_N_STREAMS * d + grp[z].x * d;
            const SplitCtx sg = split_ctx(h, grp[z].lo, chain, side);   // the group's own rows of the scratch arena
            e1.split = h->fp32_tc ? &sg : nullptr; e1.w_static = 1; e1.a_from_ln = 4;
            CFB_TRY(gemm(a_abs + (size_t)grp[z].lo * d, 0, d, (const float*)w.w_qx + (size_t)grp[z].x * d * d, 0, d,
                         grp[z].rows, d, d, 0, e1, sc));
          }
        }
        ca.skip_slot0 = 1;
        CFB_TRY(cross_attention<T>(qx_abs, h->mem_hat.as<T>(), uc, ca, n_batch, h->ntok, d, sc));
        if (side) CFB_CUDA(cudaEventRecord(aux->ev_b, sc));
      }
      auto cond_fuser = [&]() -> int {
        if (ng == 0) return CFB_OK;
        if (side) CFB_CUDA(cudaStreamWaitEvent(st, aux->ev_b, 0));
        Epilogue eg{}; eg.bias_period = 1; eg.accumulate = 1; eg.ldo = d; eg.replicate = 1;
        eg.ab_f16 = uc_f16;
        const bf16* wfu = uc_f16 ? w16.w_fu : (const bf16*)w.w_fu;
        if constexpr (sizeof(T) == 2) {
          for (int r = 0; r < sp->n_rounds; ++r) {
            TcGroup round[TC_MAX_GROUPS];
            int n = 0;
            for (int z = 0; z < ng; ++z)
              if (grp[z].round == r)
                round[n++] = TcGroup{uc + grp[z].x * d, wfu + grp[z].x * d, nullptr, h_abs, grp[z].lo, grp[z].rows};
            if (n > 0) CFB_TRY(gemm_tc_grouped(round, n, R_total, CFB_N_STREAMS * d, CFB_N_STREAMS * d, d, d, eg, st));
          }
        } else {
          for (int z = 0; z < ng; ++z) {   // one stream: sequential accumulation, overlapping row blocks are fine
            Epilogue e1 = eg; e1.out = h_abs + (size_t)grp[z].lo * d;
            const SplitCtx sg = split_ctx(h, grp[z].lo, chain, false);
            e1.split = h->fp32_tc ? &sg : nullptr; e1.w_static = 1; e1.a_from_ln = 128;
            CFB_TRY(gemm(uc + (size_t)grp[z].lo * CFB_N_STREAMS * d + grp[z].x * d, 0, CFB_N_STREAMS * d,
                         (const float*)w.w_fu + grp[z].x * d, 0, CFB_N_STREAMS * d, grp[z].rows, d, d, 0, e1, st));
          }
        }
        return CFB_OK;
      };
      // pairs on slot 0: scores against the pre-projected keys, softmax, pre-projected values straight into h
      float* sS = h->sS.as<float>() + (size_t)row0 * sp->n_tot;
      T* sP = h->sP.as<T>() + (size_t)row0 * sp->k_tot;
      Epilogue es{}; es.bias = h->z0all.as<float>() + (size_t)l * sp->n_tot; es.bias_period = 1; es.out = sS;
      es.ldo = sp->n_tot; es.replicate = 1; es.split = scm; es.a_from_ln = 4;   // keys Z change every step: this chain's W slot
      es.ab_f16 = (F16_SITES & 4) ? 1 : 0;                                          // fp16 norm2 output x fp16 keys (shared_precompute)
      CFB_TRY(gemm(a, tb, d, h->zall.as<T>() + (size_t)l * d, tb, Ld, R, sp->n_tot, d, 0, es, st));
      SharedAttnArgs sa{};
      for (int x = 0; x < CFB_N_STREAMS; ++x) {
        sa.len[x] = ca.len[x]; sa.s_off[x] = sp->s_off[x]; sa.p_off[x] = sp->p_off[x]; sa.kp[x] = sp->kp[x];
        sa.slot[x] = ca.slot[x]; sa.mask[x] = ca.mask[x];
      }
      sa.ld_s = sp->n_tot; sa.ld_p = sp->k_tot; sa.bs_offset = b0; sa.p_f16 = pv_f16;
      CFB_TRY(softmax_shared<T>(sS, sP, sa, n_batch, h->ntok, st));
      if (rb & 2) {   // fuser block + shared values -> time_block2 -> norm3 on resident rows (rowblock.cu)
        if (side) CFB_CUDA(cudaStreamWaitEvent(st, aux->ev_b, 0));
        CFB_TRY(rowblock_run(h, l, 1, row0, R, R_total, step_ptr, st));
        rb2_done = true;
      } else {
        Epilogue ey{}; ey.bias = w.b_fu; ey.bias_period = 1; ey.accumulate = 1; ey.out = hres; ey.ldo = d; ey.replicate = 1;
        ey.split = scm; ey.a_from_ln = 256; ey.ab_f16 = pv_f16;   // fp16 probabilities x fp16 values (shared_precompute)
        CFB_TRY(gemm(sP, tb, sp->k_tot, h->ytall.as<T>() + (size_t)l * d * sp->k_tot, tb, sp->k_tot, R, d, sp->k_tot, 0, ey, st));
        CFB_TRY(cond_fuser());
        CFB_TRY(ln_to(2, w.tb2_g, w.tb2_b, mod2));
      }
      shared_done = true;
    }
    if (!shared_done) {
      CFB_TRY(lin_T(d, w.w_qx, w16.w_qx, w.b_qx, qx, CFB_N_STREAMS * d, 0, 4, pr_f16));
      CFB_TRY(cross_attention<T>(qx_abs, h->mem_hat.as<T>(), qx_abs, ca, n_batch, h->ntok, d, st));
      CFB_TRY(lin_res_ln(qx, CFB_N_STREAMS * d, w.w_fu, w16.w_fu, w.b_fu, w.tb2_g, w.tb2_b, mod2, 128, 2));   // + time_block2 prologue (:655)
    }
    if (!rb2_done) CFB_TRY(lin_res_ln(a, d, w.w_tb2, w16.w_tb2, w.b_tb2, w.ln3_g, w.ln3_b, nullptr, 2, 8));   // + norm3 (:659)
    // feed-forward (:659-661); the update carries the next layer's norm1 (or the final decoder.norm)
    CFB_TRY(lin_T(d, w.w_ff1, w16.w_ff1, w.b_ff1, f, h->ff, CFB_ACT_GELU, 8, (F16_SITES & 64) ? 1 : 0));
    const bool last = l + 1 == h->L;
    if ((rb & 4) && !(last && (sites & 16))) {   // linear2 -> next norm1 on resident rows (rowblock.cu)
      CFB_TRY(rowblock_run(h, l, 2, row0, R, R_total, step_ptr, st));
    } else {
      CFB_TRY(lin_res_ln(f, h->ff, w.w_ff2, w16.w_ff2, w.b_ff2, last ? h->w.lnf_g : h->layers[l + 1].ln1_g,
                         last ? h->w.lnf_b : h->layers[l + 1].ln1_b, nullptr, 64, last ? 16 : 1));
    }
  }
  // latent_proj on a = decoder.norm(h) (cross_attention.py:238-239, denoiser.py:382)
  Epilogue ep{}; ep.bias = h->w.b_out; ep.bias_period = 1; ep.out = eps_out; ep.ldo = h->lat; ep.replicate = 1;
  ep.split = scm; ep.w_static = 1; ep.a_from_ln = 16;
  const int at = (sites & 16) ? 2 : 1;
  ep.a_terms = at; ep.ab_f16 = (F16_SITES & 16) ? 1 : 0;
  return gemm(at == 2 ? (const void*)a2 : (const void*)a, tb, at * d, ep.ab_f16 ? (const void*)h->w_out16 : h->w.w_out, tb, d,
              R, h->lat, d, 0, ep, st);
}

// time embedding + TimeBlock (scale|shift) tables for S timesteps already in h->tsteps (float)
int prep_time(cfb_denoiser* h, int S, cudaStream_t st) {
  const int d = h->d;
  const int nmod = h->L * 2 * 2 * d;
  CFB_TRY(h->tsin.reserve((size_t)S * d * 4, &h->epoch));
  CFB_TRY(h->t1.reserve((size_t)S * d * 4, &h->epoch));
  CFB_TRY(h->temb.reserve((size_t)S * d * 4, &h->epoch));
  CFB_TRY(h->tbmod.reserve((size_t)S * nmod * 4, &h->epoch));
  CFB_TRY(time_sinusoid(h->tsteps.as<float>(), h->tsin.as<float>(), S, d, st));   // denoiser.py:195-196
  Epilogue e1{}; e1.bias = h->w.b_t1; e1.bias_period = 1; e1.act = CFB_ACT_SILU; e1.out = h->t1.p; e1.ldo = d; e1.replicate = 1;
  CFB_TRY(gemm_simt(h->tsin.p, 0, d, h->w.w_t1, 0, d, S, d, d, 0, e1, st));       // embeddings.py:298-305
  Epilogue e2{}; e2.bias = h->w.b_t2; e2.bias_period = 1; e2.out = h->temb.p; e2.ldo = d; e2.replicate = 1;
  CFB_TRY(gemm_simt(h->t1.p, 0, d, h->w.w_t2, 0, d, S, d, d, 0, e2, st));
  Epilogue e3{}; e3.bias = h->w.b_tbmod; e3.bias_period = 1; e3.out = h->tbmod.p; e3.ldo = nmod; e3.replicate = 1;
  return gemm_simt(h->temb.p, 0, d, h->w.w_tbmod, 0, d, S, nmod, d, CFB_ACT_SILU, e3, st);  // cross_attention.py:433
}

int reserve_rows(cfb_denoiser* h, int n_batch, int n_in) {
  const size_t R = (size_t)n_batch * h->ntok, es = h->prec == CFB_BF16 ? 2 : 4, d = h->d;
  // (the two-term GEMM variant exists for the TMA-epilogue kernel only: plain operands with CFB_TC_TMA_EPI=0)
  h->act_sites = (h->prec == CFB_BF16 && g_gemm_backend != CFB_GEMM_SIMT && gemm_tc_two_term_ok()) ? (g_bf16_act_sites & 27) : 0;
  // (the row-block programs write their LayerNorm outputs as bf16)
  h->act_f16 = (h->prec == CFB_BF16 && g_gemm_backend != CFB_GEMM_SIMT && g_rowblock == 0) ? (g_bf16_act_f16 & 31) : 0;
  CFB_TRY(h->h.reserve(R * d * 4, &h->epoch));
  CFB_TRY(h->a.reserve(R * d * es, &h->epoch));
  if (h->act_sites) CFB_TRY(h->a2.reserve(R * d * 2 * 2, &h->epoch));
  CFB_TRY(h->qkv.reserve(R * 3 * d * es, &h->epoch));
  CFB_TRY(h->qx.reserve(R * CFB_N_STREAMS * d * es, &h->epoch));
  CFB_TRY(h->f.reserve(R * h->ff * es, &h->epoch));
  CFB_TRY(h->xin.reserve((size_t)n_in * h->ntok * h->lat * es, &h->epoch));
  CFB_TRY(h->eps.reserve(R * h->lat * 4, &h->epoch));
  return CFB_OK;
}

// Validate the memory description, build mem_c, stage slots/masks in handle-owned buffers.
int prep_memory(cfb_denoiser* h, const cfb_memory* mem, int n_batch, MemLayout* ml, CrossArgs* ca, cudaStream_t st) {
  int rows = 0, slot_elems = 0, mask_bytes = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    CFB_CHECK(mem->cond[x] != nullptr && mem->n_slots[x] > 0 && mem->len[x] > 0, "memory stream %d is empty", x);
    CFB_CHECK(mem->len[x] <= h->w.pe_len, "memory stream %d: %d tokens exceed the positional table (%d)", x, mem->len[x], h->w.pe_len);
    CFB_CHECK(mem->slot[x] != nullptr || mem->n_slots[x] == n_batch, "memory stream %d: %d slots for %d batch entries and no slot table", x, mem->n_slots[x], n_batch);
    ml->row_base[x] = rows; rows += mem->n_slots[x] * mem->len[x];
    ml->slot_off[x] = slot_elems; slot_elems += n_batch;
    ml->mask_off[x] = mask_bytes; mask_bytes += (mem->n_slots[x] * mem->len[x] + 15) & ~15;
  }
  ml->total_rows = rows;
  const size_t es = h->prec == CFB_BF16 ? 2 : 4;
  CFB_TRY(h->mem_c.reserve((size_t)rows * h->d * 4, &h->epoch));
  CFB_TRY(h->mem_hat.reserve((size_t)rows * h->d * es, &h->epoch));
  CFB_TRY(h->slots.reserve((size_t)slot_elems * 4, &h->epoch));
  CFB_TRY(h->masks.reserve((size_t)mask_bytes, &h->epoch));
  CFB_TRY(mem_build(mem->cond, mem->n_slots, mem->len, h->w.stream_emb, h->w.pe_mem, h->mem_c.as<float>(), h->d, st));
  memset(ca, 0, sizeof(*ca));
  long long t_elems = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    ca->n_slots[x] = mem->n_slots[x];
    ca->lenp[x] = (mem->len[x] + 7) & ~7;
    ca->t_off[x] = t_elems;
    t_elems += (long long)mem->n_slots[x] * h->d * ca->lenp[x];
    t_elems = (t_elems + 63) & ~63LL;                      // tensor maps want 128-byte aligned bases
  }
  if (h->prec == CFB_BF16) {                               // transposed copy for the tcgen05 per-pair attention
    CFB_TRY(h->mem_hat_t.reserve((size_t)t_elems * 2, &h->epoch));
    ca->mem_hat_t = h->mem_hat_t.as<bf16>();
  }
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    ca->row_base[x] = ml->row_base[x];
    ca->len[x] = mem->len[x];
    if (mem->slot[x]) {
      int* dst = h->slots.as<int>() + ml->slot_off[x];
      CFB_CUDA(cudaMemcpyAsync(dst, mem->slot[x], (size_t)n_batch * 4, cudaMemcpyDeviceToDevice, st));
      ca->slot[x] = dst;
    }
    if (mem->mask[x]) {
      uint8_t* dst = h->masks.as<uint8_t>() + ml->mask_off[x];
      CFB_CUDA(cudaMemcpyAsync(dst, mem->mask[x], (size_t)mem->n_slots[x] * mem->len[x], cudaMemcpyDeviceToDevice, st));
      ca->mask[x] = dst;
    }
  }
  return CFB_OK;
}

template <typename T>
int step_body(cfb_denoiser* h, int n_clips, int n_branch, const MemLayout& ml, const CrossArgs& ca,
              float* const att_base[CFB_N_STREAMS], const StepArgs& sa, const SharedPlan& sp, cudaStream_t st) {
  const int* step_ptr = h->step.as<int>();
  const bool overlap = sp.on && h->n_chains > 1 && h->pre_st[0] != nullptr;
  if (overlap) {
    // The per-step memory normalisation and the keys / values of the shared slot are needed first by layer 0's
    // cross-attention, so they leave the critical path entirely: two side streams, joined by ev_pre in run_layers.
    CFB_CUDA(cudaEventRecord(h->ev_fork, st));
    CFB_CUDA(cudaStreamWaitEvent(h->pre_st[0], h->ev_fork, 0));
    CFB_TRY(mem_hat<T>(h->mem_c.as<float>(), h->temb.as<float>(), step_ptr, h->mem_hat.as<T>(), ml.total_rows, h->d, h->pre_st[0], pair_f16(h)));
    CFB_CUDA(cudaEventRecord(h->ev_mh, h->pre_st[0]));
    if constexpr (sizeof(T) == 2)
      if (cross_tc_supported(ca, h->ntok, h->d)) CFB_TRY(mem_transpose(h->mem_hat.as<bf16>(), h->mem_hat_t.as<bf16>(), ca, h->pre_st[0]));
    CFB_CUDA(cudaStreamWaitEvent(h->pre_st[1], h->ev_mh, 0));
    for (int i = 0; i < 2; ++i) {
      CFB_TRY(shared_precompute<T>(h, sp, ml, ca.len, i, h->pre_st[i]));
      CFB_CUDA(cudaEventRecord(h->ev_pre[i], h->pre_st[i]));
    }
  } else {
    CFB_TRY(mem_hat<T>(h->mem_c.as<float>(), h->temb.as<float>(), step_ptr, h->mem_hat.as<T>(), ml.total_rows, h->d, st, pair_f16(h)));
    if constexpr (sizeof(T) == 2)
      if (cross_tc_supported(ca, h->ntok, h->d)) CFB_TRY(mem_transpose(h->mem_hat.as<bf16>(), h->mem_hat_t.as<bf16>(), ca, st));
    if (sp.on) {
      CFB_TRY(shared_precompute<T>(h, sp, ml, ca.len, 0, st));
      CFB_TRY(shared_precompute<T>(h, sp, ml, ca.len, 1, st));
    }
  }
  CFB_TRY(embed_cast<T>(h, h->x.as<float>(), n_clips, st));       // torch.cat([latents] * 7), convofusion.py:499
  // The step is a chain of ~190 short kernels, bound by launch / prologue / epilogue latency rather than by
  // throughput.  Batch entries are independent through the whole denoiser, so they are cut into n_chains groups
  // (multiples of 8 entries = one 128-row tile) that run the layer stack concurrently on forked streams.
  const int n_batch = n_clips * n_branch;
  const int per = ((n_batch + h->n_chains - 1) / h->n_chains + 7) & ~7;
  const int nc = (n_batch + per - 1) / per;
  if (nc <= 1) {
    ChainAux aux;
    if (overlap) { aux.ev_pre[0] = h->ev_pre[0]; aux.ev_pre[1] = h->ev_pre[1]; }
    CFB_TRY(embed_rows<T>(h, 0, n_batch, n_clips, st));
    CFB_TRY(run_layers<T>(h, n_batch, ca, att_base, step_ptr, h->eps.as<float>(), st, &sp, 0, n_batch, &aux, true));
  } else {
    CFB_CUDA(cudaEventRecord(h->ev_fork, st));
    for (int c = 0; c < nc; ++c) {
      const int b0 = c * per, nb = (b0 + per <= n_batch ? per : n_batch - b0);
      cudaStream_t cs = c == 0 ? st : h->chain_st[c];
      if (c > 0) CFB_CUDA(cudaStreamWaitEvent(cs, h->ev_fork, 0));
      ChainAux aux;
      if (overlap) {
        aux.ev_pre[0] = h->ev_pre[0]; aux.ev_pre[1] = h->ev_pre[1];
        aux.st2 = h->chain_st2[c]; aux.ev_a = h->ev_a[c]; aux.ev_b = h->ev_b[c];
      }
      CFB_TRY(embed_rows<T>(h, b0, nb, n_clips, cs));
      CFB_TRY(run_layers<T>(h, nb, ca, att_base, step_ptr, h->eps.as<float>(), cs, &sp, b0, n_batch, &aux, true, c));
      if (c > 0) {
        CFB_CUDA(cudaEventRecord(h->ev_join[c], cs));
        CFB_CUDA(cudaStreamWaitEvent(st, h->ev_join[c], 0));
      }
    }
  }
  return guidance_sched_step(sa, st);
}

// Decide whether the shared-slot plan applies and derive the conditional row groups from the slot tables.  The plan
// runs in both precisions (fp32: the same algebra on the CUDA-core GEMM, which is how the plan is checked against the
// oracle at 1e-4); env CFB_PLAN=0 keeps the general per-pair path.  The tables are read from mem->slot_host when the
// caller supplies a host copy (no device read-back, no stream synchronisation).
int make_shared_plan(cfb_denoiser* h, const cfb_memory* mem, int n_batch, SharedPlan* sp, cudaStream_t st) {
  sp->on = false;
  if (!g_shared_plan || (h->prec == CFB_BF16 && g_gemm_backend == CFB_GEMM_SIMT)) return CFB_OK;
  for (int x = 0; x < CFB_N_STREAMS; ++x)
    if (mem->slot[x] == nullptr) return CFB_OK;
  std::vector<int> hs((size_t)n_batch);
  int ng = 0, s_off = 0, p_off = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    const int* tab = mem->slot_host[x];
    if (tab == nullptr) {
      CFB_CUDA(cudaMemcpyAsync(hs.data(), mem->slot[x], (size_t)n_batch * 4, cudaMemcpyDeviceToHost, st));
      CFB_CUDA(cudaStreamSynchronize(st));
      tab = hs.data();
    }
    for (int b = 0; b < n_batch;) {
      CFB_CHECK(tab[b] >= 0 && tab[b] < mem->n_slots[x], "memory stream %d: slot index %d out of range", x, tab[b]);
      if (tab[b] == 0) { ++b; continue; }
      int e = b;
      while (e < n_batch && tab[e] != 0) {
        CFB_CHECK(tab[e] >= 0 && tab[e] < mem->n_slots[x], "memory stream %d: slot index %d out of range", x, tab[e]);
        ++e;
      }
      if (ng == TC_MAX_GROUPS) return CFB_OK;   // too fragmented: keep the general path
      sp->g_stream[ng] = x; sp->g_row_start[ng] = b * h->ntok; sp->g_rows[ng] = (e - b) * h->ntok;
      ++ng;
      b = e;
    }
    sp->s_off[x] = s_off; s_off += (mem->len[x] + 31) & ~31;
    sp->kp[x] = (mem->len[x] + 63) & ~63;
    sp->p_off[x] = p_off; p_off += sp->kp[x];
  }
  sp->n_tot = s_off; sp->k_tot = p_off; sp->n_groups = ng;
  sp->n_rounds = 0;
  for (int z = 0; z < ng; ++z) {   // greedy interval colouring
    bool used[TC_MAX_GROUPS] = {};
    for (int y = 0; y < z; ++y) {
      const bool overlap = sp->g_row_start[y] < sp->g_row_start[z] + sp->g_rows[z] &&
                           sp->g_row_start[z] < sp->g_row_start[y] + sp->g_rows[y];
      if (overlap) used[sp->g_round[y]] = true;
    }
    int r = 0;
    while (used[r]) ++r;
    sp->g_round[z] = r;
    if (r + 1 > sp->n_rounds) sp->n_rounds = r + 1;
  }
  const size_t R = (size_t)n_batch * h->ntok, Ld = (size_t)h->L * h->d, es = h->prec == CFB_BF16 ? 2 : 4;
  const unsigned before = h->epoch;
  CFB_TRY(h->uc.reserve(R * CFB_N_STREAMS * h->d * es, &h->epoch));
  CFB_TRY(h->sS.reserve(R * sp->n_tot * 4, &h->epoch));
  CFB_TRY(h->sP.reserve(R * sp->k_tot * es, &h->epoch));
  CFB_TRY(h->zall.reserve((size_t)sp->n_tot * Ld * es, &h->epoch));
  CFB_TRY(h->z0all.reserve((size_t)h->L * sp->n_tot * 4, &h->epoch));
  CFB_TRY(h->ytall.reserve(Ld * sp->k_tot * es, &h->epoch));
  if (h->epoch != before) {   // fresh allocations: padding rows/columns must hold finite values
    CFB_CUDA(cudaMemsetAsync(h->zall.p, 0, h->zall.cap, st));
    CFB_CUDA(cudaMemsetAsync(h->z0all.p, 0, h->z0all.cap, st));
    CFB_CUDA(cudaMemsetAsync(h->ytall.p, 0, h->ytall.cap, st));
    CFB_CUDA(cudaMemsetAsync(h->uc.p, 0, h->uc.cap, st));
  }
  sp->on = true;
  return CFB_OK;
}

}  // namespace

extern "C" {

int cfb_denoiser_create(const cfb_denoiser_weights* w, cfb_denoiser** out) {
  CFB_CHECK(w && out, "cfb_denoiser_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: convofusion_b200 has no CPU fallback");
    return CFB_ERR_NO_DEVICE;
  }
  CFB_CHECK(w->d_model == 512, "d_model %d unsupported (512)", w->d_model);
  CFB_CHECK(w->n_tokens > 0 && w->n_tokens <= 16, "n_tokens %d unsupported (<=16)", w->n_tokens);
  CFB_CHECK(w->latent_dim % 64 == 0 && w->ff_size % 64 == 0, "latent_dim/ff_size must be multiples of 64");
  CFB_CHECK(w->precision == CFB_F32 || w->precision == CFB_BF16, "bad precision");
  CFB_CHECK(w->n_layers > 0 && w->layers != nullptr, "no layers");
  cfb_denoiser* h = new cfb_denoiser();
  h->w = *w;
  h->layers.assign(w->layers, w->layers + w->n_layers);
  h->w.layers = h->layers.data();
  h->d = w->d_model; h->lat = w->latent_dim; h->ntok = w->n_tokens; h->L = w->n_layers; h->H = w->n_heads;
  h->ff = w->ff_size; h->prec = w->precision;
  int rc = init_gemm_tc_kernels();
  if (rc == CFB_OK) rc = init_attention_kernels();
  if (rc == CFB_OK) rc = init_cross_tc_kernels();
  // fp16 copies of the weights fed by LayerNorm outputs (g_bf16_act_f16): qkv, TimeBlock linears, linear1, latent_proj
  if (rc == CFB_OK && h->prec == CFB_BF16) {
    auto make16 = [&]() -> int {
      const size_t d2 = (size_t)h->d * h->d, ffd = (size_t)h->ff * h->d, per = (6 + 2 * CFB_N_STREAMS) * d2 + 2 * ffd;
      const size_t outn = (size_t)h->lat * h->d;
      const size_t pre = (size_t)h->L * d2;
      CFB_TRY(h->w16.reserve((per * h->L + 2 * outn + 2 * CFB_N_STREAMS * pre) * 2, nullptr));
      bf16* p = h->w16.as<bf16>();
      auto conv = [&](const void* src, size_t n, const bf16** dst) -> int {
        CFB_TRY(bf16_to_f16((const bf16*)src, p, n, nullptr));
        *dst = p; p += n;
        return CFB_OK;
      };
      h->l16.resize(h->L);
      for (int l = 0; l < h->L; ++l) {
        CFB_TRY(conv(h->layers[l].w_in, 3 * d2, &h->l16[l].w_in));
        CFB_TRY(conv(h->layers[l].w_tb1, d2, &h->l16[l].w_tb1));
        CFB_TRY(conv(h->layers[l].w_tb2, d2, &h->l16[l].w_tb2));
        CFB_TRY(conv(h->layers[l].w_ff1, ffd, &h->l16[l].w_ff1));
        CFB_TRY(conv(h->layers[l].w_fu, CFB_N_STREAMS * d2, &h->l16[l].w_fu));
        CFB_TRY(conv(h->layers[l].w_qx, CFB_N_STREAMS * d2, &h->l16[l].w_qx));
        CFB_TRY(conv(h->layers[l].w_so, d2, &h->l16[l].w_so));
        CFB_TRY(conv(h->layers[l].w_ff2, ffd, &h->l16[l].w_ff2));
      }
      CFB_TRY(conv(h->w.w_out, outn, &h->w_out16));
      CFB_TRY(conv(h->w.w_embed, outn, &h->w_embed16));
      for (int x = 0; x < CFB_N_STREAMS; ++x) {
        CFB_TRY(conv(h->w.w_zx[x], pre, &h->w_zx16[x]));
        CFB_TRY(conv(h->w.w_yx[x], pre, &h->w_yx16[x]));
      }
      CFB_CUDA(cudaStreamSynchronize(nullptr));
      return CFB_OK;
    };
    rc = make16();
  }
  if (rc != CFB_OK) { cfb_denoiser_destroy(h); return rc; }
  *out = h;
  return CFB_OK;
}

void cfb_denoiser_destroy(cfb_denoiser* h) {
  if (!h) return;
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  if (h->ev_out) cudaEventDestroy(h->ev_out);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int c = 0; c < cfb_denoiser::MAX_CHAINS; ++c) {
    if (h->chain_st[c]) cudaStreamDestroy(h->chain_st[c]);
    if (h->chain_st2[c]) cudaStreamDestroy(h->chain_st2[c]);
    if (h->ev_join[c]) cudaEventDestroy(h->ev_join[c]);
    if (h->ev_a[c]) cudaEventDestroy(h->ev_a[c]);
    if (h->ev_b[c]) cudaEventDestroy(h->ev_b[c]);
  }
  for (int i = 0; i < 2; ++i) {
    if (h->pre_st[i]) cudaStreamDestroy(h->pre_st[i]);
    if (h->ev_pre[i]) cudaEventDestroy(h->ev_pre[i]);
  }
  if (h->ev_mh) cudaEventDestroy(h->ev_mh);
  if (h->ev_sched) cudaEventDestroy(h->ev_sched);
  DeviceBuf* bufs[] = {&h->h, &h->a, &h->qkv, &h->qx, &h->f, &h->xin, &h->eps, &h->mem_c, &h->mem_hat, &h->tsteps,
                       &h->tsin, &h->t1, &h->temb, &h->tbmod, &h->coef, &h->step, &h->x, &h->inp_noise, &h->preseq,
                       &h->slots, &h->masks, &h->uc, &h->sS, &h->sP, &h->zall, &h->z0all, &h->ytall, &h->rb_prog, &h->rb_blk, &h->split_ws,
                       &h->wg_h, &h->wg_qkv, &h->wg_z, &h->wg_p, &h->wg_g, &h->wg_t1, &h->wg_t2, &h->mem_hat_t, &h->a2, &h->w16};
  for (DeviceBuf* b : bufs) b->release();
  split_cache_destroy(h->split_cache);
  delete h;
}

int cfb_set_rowblock(int mask) {
  g_rowblock = mask & 7;
  return CFB_OK;
}

int cfb_set_fp32_tensor_cores(int mode) {
  CFB_CHECK(mode >= 0 && mode <= 4, "cfb_set_fp32_tensor_cores: mode %d outside 0..4", mode);
  g_fp32_tc = mode;
  return CFB_OK;
}

int cfb_set_bf16_activation_terms(int terms) {   // 2: every site, 1: none
  CFB_CHECK(terms == 1 || terms == 2, "cfb_set_bf16_activation_terms: %d (1 or 2)", terms);
  g_bf16_act_sites = terms == 2 ? 27 : 0;
  return CFB_OK;
}

int cfb_denoiser_attach_f16_weights(cfb_denoiser* h, const cfb_denoiser_weights* w16) {
  CFB_CHECK(h && w16 && w16->layers, "cfb_denoiser_attach_f16_weights: null argument");
  CFB_CHECK(h->prec == CFB_BF16, "cfb_denoiser_attach_f16_weights: the handle is not a 16-bit handle");
  CFB_CHECK(w16->n_layers == h->L && w16->d_model == h->d && w16->ff_size == h->ff && w16->latent_dim == h->lat,
            "cfb_denoiser_attach_f16_weights: shape mismatch");
  h->l16.resize(h->L);
  for (int l = 0; l < h->L; ++l) {
    const cfb_denoiser_layer& s = w16->layers[l];
    CFB_CHECK(s.w_in && s.w_so && s.w_tb1 && s.w_tb2 && s.w_qx && s.w_fu && s.w_ff1 && s.w_ff2,
              "cfb_denoiser_attach_f16_weights: layer %d lacks a matrix", l);
    h->l16[l] = cfb_denoiser::W16{(const bf16*)s.w_in, (const bf16*)s.w_tb1, (const bf16*)s.w_tb2, (const bf16*)s.w_ff1,
                                  (const bf16*)s.w_fu, (const bf16*)s.w_qx, (const bf16*)s.w_so, (const bf16*)s.w_ff2};
  }
  CFB_CHECK(w16->w_out && w16->w_embed, "cfb_denoiser_attach_f16_weights: latent_proj / latent_embd missing");
  h->w_out16 = (const bf16*)w16->w_out;
  h->w_embed16 = (const bf16*)w16->w_embed;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    CFB_CHECK(w16->w_zx[x] && w16->w_yx[x], "cfb_denoiser_attach_f16_weights: pre-projection %d missing", x);
    h->w_zx16[x] = (const bf16*)w16->w_zx[x]; h->w_yx16[x] = (const bf16*)w16->w_yx[x];
  }
  h->w16.release();   // (cudaFree synchronises: the conversions of cfb_denoiser_create have long finished)
  if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }   // captured pointers are stale
  h->graph_valid = false;
  return CFB_OK;
}

int cfb_set_bf16_activation_f16(int mask) {
  CFB_CHECK(mask >= 0 && mask <= 31, "cfb_set_bf16_activation_f16: mask %d outside 0..31", mask);
  g_bf16_act_f16 = mask;
  return CFB_OK;
}

int cfb_set_bf16_activation_sites(int sites) {
  CFB_CHECK(sites >= 0 && (sites & ~27) == 0, "cfb_set_bf16_activation_sites: mask %d outside {1, 2, 8, 16}", sites);
  g_bf16_act_sites = sites;
  return CFB_OK;
}

int cfb_set_cross_tc(int enabled) {
  g_cross_tc = enabled != 0;
  return CFB_OK;
}

int cfb_set_shared_plan(int enabled) {
  g_shared_plan = enabled != 0;
  return CFB_OK;
}

int cfb_denoiser_set_chains(cfb_denoiser* h, int n_chains) {
  CFB_CHECK(h != nullptr, "cfb_denoiser_set_chains: null handle");
  CFB_CHECK(n_chains >= 0 && n_chains <= cfb_denoiser::MAX_CHAINS, "n_chains %d outside 0..%d", n_chains, cfb_denoiser::MAX_CHAINS);
  h->chains_override = n_chains;
  return CFB_OK;
}

int cfb_denoiser_forward(cfb_denoiser* h, const float* sample, int n_batch, int64_t timestep, const cfb_memory* mem,
                         float* eps_out, float* const att_out[CFB_N_STREAMS], cfb_stream stream) {
  CFB_CHECK(h && sample && mem && eps_out && n_batch > 0, "cfb_denoiser_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  h->weg.valid = false;     // the workspace the saved activations refer to is about to be reused
  CFB_TRY(reserve_rows(h, n_batch, n_batch));
  h->sched_epoch = ~0u;   // the single-step tables below replace those of the last cfb_sample call
  CFB_TRY(h->tsteps.reserve(4, &h->epoch));
  const float tf = (float)timestep;
  CFB_CUDA(cudaMemcpyAsync(h->tsteps.p, &tf, 4, cudaMemcpyHostToDevice, st));
  CFB_TRY(prep_time(h, 1, st));
  MemLayout ml; CrossArgs ca;
  CFB_TRY(prep_memory(h, mem, n_batch, &ml, &ca, st));
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    ca.att_batch_stride[x] = (long long)h->L * h->ntok * mem->len[x];
    ca.att_step_stride[x] = 0;
  }
  ca.att_first_batch = 0; ca.step_ptr = nullptr;
  CFB_TRY(reserve_split(h, n_batch, mem, false));
  if (h->prec == CFB_BF16) {
    CFB_TRY(mem_hat<bf16>(h->mem_c.as<float>(), h->temb.as<float>(), nullptr, h->mem_hat.as<bf16>(), ml.total_rows, h->d, st, pair_f16(h)));
    if (cross_tc_supported(ca, h->ntok, h->d)) CFB_TRY(mem_transpose(h->mem_hat.as<bf16>(), h->mem_hat_t.as<bf16>(), ca, st));
    CFB_TRY(embed<bf16>(h, sample, n_batch, 1, st));
    return run_layers<bf16>(h, n_batch, ca, att_out, nullptr, eps_out, st);
  }
  CFB_TRY(mem_hat<float>(h->mem_c.as<float>(), h->temb.as<float>(), nullptr, h->mem_hat.as<float>(), ml.total_rows, h->d, st));
  CFB_TRY(embed<float>(h, sample, n_batch, 1, st));
  return run_layers<float>(h, n_batch, ca, att_out, nullptr, eps_out, st);
}


// ---- word-excitation guidance (convofusion.py:437-496, 298-388) ---------------------------------------------------
// The reference differentiates a loss on the listener-text attention maps of the text-only branch with respect to the
// latents (torch.autograd through Denoiser.forward).  Here: one fp32 evaluation that keeps what the backward needs
// (the five LayerNorm inputs of every layer, qkv, the pre-GELU activations, all cross-attention probabilities), then
// the reverse walk with the input-gradient kernels of weg.cu.  fp32 handles only; batch = the text-only branch.
int cfb_denoiser_weg_forward(cfb_denoiser* h, const float* sample, int n_batch, int64_t timestep, const cfb_memory* mem,
                             int att_stream, float* att_out, cfb_stream stream) {
  CFB_CHECK(h && sample && mem && att_out && n_batch > 0, "cfb_denoiser_weg_forward: bad argument");
  CFB_CHECK(h->prec == CFB_F32, "cfb_denoiser_weg_forward: word-excitation guidance runs on an fp32 handle");
  CFB_CHECK(att_stream >= 0 && att_stream < CFB_N_STREAMS, "cfb_denoiser_weg_forward: bad stream %d", att_stream);
  cudaStream_t st = (cudaStream_t)stream;
  h->weg.valid = false;
  const int d = h->d, L = h->L, R = n_batch * h->ntok, ff = h->ff;
  CFB_TRY(reserve_rows(h, n_batch, n_batch));
  h->sched_epoch = ~0u;
  CFB_TRY(h->tsteps.reserve(4, &h->epoch));
  const float tf = (float)timestep;
  CFB_CUDA(cudaMemcpyAsync(h->tsteps.p, &tf, 4, cudaMemcpyHostToDevice, st));
  CFB_TRY(prep_time(h, 1, st));
  MemLayout ml; CrossArgs ca;
  CFB_TRY(prep_memory(h, mem, n_batch, &ml, &ca, st));
  long long p_total = 0;
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    h->weg.p_off[x] = p_total;
    p_total += (long long)n_batch * L * h->ntok * mem->len[x];
    ca.att_batch_stride[x] = (long long)L * h->ntok * mem->len[x];
    ca.att_step_stride[x] = 0;
  }
  ca.att_first_batch = 0; ca.step_ptr = nullptr; ca.skip_slot0 = 0; ca.bs_offset = 0;
  const size_t Rd = (size_t)R * d;
  int wide = CFB_N_STREAMS * d;
  if (ff > wide) wide = ff;
  CFB_TRY(h->wg_h.reserve((size_t)L * 5 * Rd * 4, nullptr));
  CFB_TRY(h->wg_qkv.reserve((size_t)L * Rd * 3 * 4, nullptr));
  CFB_TRY(h->wg_z.reserve((size_t)L * R * ff * 4, nullptr));
  CFB_TRY(h->wg_p.reserve((size_t)p_total * 4, nullptr));
  CFB_TRY(h->wg_g.reserve(Rd * 4, nullptr));
  CFB_TRY(h->wg_t1.reserve((size_t)R * wide * 4, nullptr));
  CFB_TRY(h->wg_t2.reserve((size_t)R * wide * 4, nullptr));
  CFB_TRY(mem_hat<float>(h->mem_c.as<float>(), h->temb.as<float>(), nullptr, h->mem_hat.as<float>(), ml.total_rows, d, st));
  CFB_TRY(embed<float>(h, sample, n_batch, 1, st));
  float* hres = h->h.as<float>();
  float* a = h->a.as<float>();
  float* qx = h->qx.as<float>();
  float* f = h->f.as<float>();
  auto snap = [&](int l, int k) {
    return cudaMemcpyAsync(h->wg_h.as<float>() + ((size_t)l * 5 + k) * Rd, hres, Rd * 4, cudaMemcpyDeviceToDevice, st);
  };
  auto lin = [&](const float* A, int K, const void* W, const float* b, float* out, int N, int act, int accumulate) {
    Epilogue ep{}; ep.bias = b; ep.bias_period = 1; ep.act = act; ep.accumulate = accumulate; ep.out = out; ep.ldo = N; ep.replicate = 1;
    return gemm(A, 0, K, W, 0, K, R, N, K, 0, ep, st);
  };
  for (int l = 0; l < L; ++l) {
    const cfb_denoiser_layer& w = h->layers[l];
    const float* mod1 = h->tbmod.as<float>() + (size_t)(2 * l) * 2 * d;
    const float* mod2 = mod1 + 2 * d;
    float* qkv = h->wg_qkv.as<float>() + (size_t)l * Rd * 3;
    float* z = h->wg_z.as<float>() + (size_t)l * R * ff;
    CFB_CUDA(snap(l, 0));
    CFB_TRY(ln_rows<float>(hres, w.ln1_g, w.ln1_b, nullptr, nullptr, 0, a, R, d, st));
    CFB_TRY(lin(a, d, w.w_in, w.b_in, qkv, 3 * d, 0, 0));
    CFB_TRY(mha<float>(qkv, 3 * d, qkv + d, qkv + 2 * d, 3 * d, a, d, n_batch, h->ntok, h->ntok, h->H, d / h->H, nullptr, st));
    CFB_TRY(lin(a, d, w.w_so, w.b_so, hres, d, 0, 1));
    CFB_CUDA(snap(l, 1));
    CFB_TRY(ln_rows<float>(hres, w.tb1_g, w.tb1_b, mod1, nullptr, 0, a, R, d, st));
    CFB_TRY(lin(a, d, w.w_tb1, w.b_tb1, hres, d, 0, 1));
    CFB_CUDA(snap(l, 2));
    CFB_TRY(ln_rows<float>(hres, w.ln2_g, w.ln2_b, nullptr, nullptr, 0, a, R, d, st));
    CFB_TRY(lin(a, d, w.w_qx, w.b_qx, qx, CFB_N_STREAMS * d, 0, 0));
    for (int x = 0; x < CFB_N_STREAMS; ++x)
      ca.att[x] = h->wg_p.as<float>() + h->weg.p_off[x] + (size_t)l * h->ntok * ca.len[x];
    CFB_TRY(cross_attention<float>(qx, h->mem_hat.as<float>(), qx, ca, n_batch, h->ntok, d, st));
    CFB_TRY(lin(qx, CFB_N_STREAMS * d, w.w_fu, w.b_fu, hres, d, 0, 1));
    CFB_CUDA(snap(l, 3));
    CFB_TRY(ln_rows<float>(hres, w.tb2_g, w.tb2_b, mod2, nullptr, 0, a, R, d, st));
    CFB_TRY(lin(a, d, w.w_tb2, w.b_tb2, hres, d, 0, 1));
    CFB_CUDA(snap(l, 4));
    CFB_TRY(ln_rows<float>(hres, w.ln3_g, w.ln3_b, nullptr, nullptr, 0, a, R, d, st));
    CFB_TRY(lin(a, d, w.w_ff1, w.b_ff1, z, ff, 0, 0));                 // pre-activation, kept for the backward
    CFB_TRY(lin(a, d, w.w_ff1, w.b_ff1, f, ff, CFB_ACT_GELU, 0));
    CFB_TRY(lin(f, ff, w.w_ff2, w.b_ff2, hres, d, 0, 1));
  }
  const long long n_att = (long long)n_batch * L * h->ntok * mem->len[att_stream];
  CFB_CUDA(cudaMemcpyAsync(att_out, h->wg_p.as<float>() + h->weg.p_off[att_stream], (size_t)n_att * 4, cudaMemcpyDeviceToDevice, st));
  h->weg.valid = true; h->weg.n_batch = n_batch; h->weg.att_stream = att_stream; h->weg.ca = ca;
  return CFB_OK;
}

int cfb_denoiser_weg_backward(cfb_denoiser* h, const float* d_att, float* grad_sample, cfb_stream stream) {
  CFB_CHECK(h && d_att && grad_sample, "cfb_denoiser_weg_backward: bad argument");
  CFB_CHECK(h->weg.valid, "cfb_denoiser_weg_backward: no saved forward (call cfb_denoiser_weg_forward on this handle first)");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_batch = h->weg.n_batch, d = h->d, L = h->L, R = n_batch * h->ntok, ff = h->ff, sx = h->weg.att_stream;
  const size_t Rd = (size_t)R * d;
  CrossArgs ca = h->weg.ca;
  float* g = h->wg_g.as<float>();
  float* t1 = h->wg_t1.as<float>();
  float* t2 = h->wg_t2.as<float>();
  CFB_CUDA(cudaMemsetAsync(g, 0, Rd * 4, st));
  auto hs = [&](int l, int k) { return h->wg_h.as<float>() + ((size_t)l * 5 + k) * Rd; };
  for (int l = L - 1; l >= 0; --l) {
    const cfb_denoiser_layer& w = h->layers[l];
    const float* mod1 = h->tbmod.as<float>() + (size_t)(2 * l) * 2 * d;
    const float* mod2 = mod1 + 2 * d;
    const float* qkv = h->wg_qkv.as<float>() + (size_t)l * Rd * 3;
    const float* z = h->wg_z.as<float>() + (size_t)l * R * ff;
    // feed-forward (cross_attention.py:659-661): h5 = h4 + linear2(GELU(linear1(norm3(h4))))
    CFB_TRY(linear_bwd(g, d, (const float*)w.w_ff2, ff, t1, ff, R, d, ff, 0, st));
    CFB_TRY(gelu_bwd(z, t1, (long long)R * ff, st));
    CFB_TRY(linear_bwd(t1, ff, (const float*)w.w_ff1, d, t2, d, R, ff, d, 0, st));
    CFB_TRY(ln_bwd(hs(l, 4), w.ln3_g, w.ln3_b, nullptr, t2, g, R, d, st));
    // time_block2 (:655)
    CFB_TRY(linear_bwd(g, d, (const float*)w.w_tb2, d, t1, d, R, d, d, 0, st));
    CFB_TRY(ln_bwd(hs(l, 3), w.tb2_g, w.tb2_b, mod2, t1, g, R, d, st));
    // five folded cross-attentions + att_fuser (:578-652); the loss enters through the maps of stream `sx`
    CFB_TRY(linear_bwd(g, d, (const float*)w.w_fu, CFB_N_STREAMS * d, t1, CFB_N_STREAMS * d, R, d, CFB_N_STREAMS * d, 0, st));
    for (int x = 0; x < CFB_N_STREAMS; ++x)
      ca.att[x] = h->wg_p.as<float>() + h->weg.p_off[x] + (size_t)l * h->ntok * ca.len[x];
    CFB_TRY(cross_bwd(t1, h->mem_hat.as<float>(), t2, ca, d_att + (size_t)l * h->ntok * ca.len[sx], sx, n_batch, h->ntok, st));
    CFB_TRY(linear_bwd(t2, CFB_N_STREAMS * d, (const float*)w.w_qx, d, t1, d, R, CFB_N_STREAMS * d, d, 0, st));
    CFB_TRY(ln_bwd(hs(l, 2), w.ln2_g, w.ln2_b, nullptr, t1, g, R, d, st));
    // time_block1 (:575)
    CFB_TRY(linear_bwd(g, d, (const float*)w.w_tb1, d, t1, d, R, d, d, 0, st));
    CFB_TRY(ln_bwd(hs(l, 1), w.tb1_g, w.tb1_b, mod1, t1, g, R, d, st));
    // self-attention (:568-572)
    CFB_TRY(linear_bwd(g, d, (const float*)w.w_so, d, t1, d, R, d, d, 0, st));
    CFB_TRY(mha_bwd(qkv, t1, t2, n_batch, h->ntok, h->H, d, st));
    CFB_TRY(linear_bwd(t2, 3 * d, (const float*)w.w_in, d, t1, d, R, 3 * d, d, 0, st));
    CFB_TRY(ln_bwd(hs(l, 0), w.ln1_g, w.ln1_b, nullptr, t1, g, R, d, st));
  }
  // latent_embd (denoiser.py:187): h = x W^T + per-token bias
  return linear_bwd(g, d, (const float*)h->w.w_embed, h->lat, grad_sample, h->lat, R, d, h->lat, 0, st);
}

int cfb_sample(cfb_denoiser* h, const cfb_schedule* sched, const cfb_memory* mem, int n_clips, int n_branch,
               int full_last, float* latents, const float* step_noise, const float* preseq, int preseq_len, float* record,
               float* const att_out[CFB_N_STREAMS], int use_graph, cfb_stream stream) {
  CFB_CHECK(h && sched && mem && latents && n_clips > 0, "cfb_sample: bad argument");
  CFB_CHECK(n_branch >= 1 && n_branch <= CFB_N_BRANCH, "cfb_sample: n_branch must be 1..7");
  CFB_CHECK(!full_last || n_branch >= 2, "cfb_sample: full_last needs at least the unconditional and the full-cond branch");
  CFB_CHECK(sched->n_steps > 0 && sched->timesteps && sched->coef, "cfb_sample: empty schedule");
  CFB_CHECK(sched->kind == CFB_SCHED_DDIM || sched->kind == CFB_SCHED_DDPM, "cfb_sample: unknown scheduler kind");
  CFB_CHECK(preseq == nullptr || (preseq_len > 0 && preseq_len <= h->ntok), "cfb_sample: bad preseq_len %d", preseq_len);
  h->weg.valid = false;     // the workspace the saved activations refer to is about to be reused
  bool want_att = false;
  if (att_out) for (int x = 0; x < CFB_N_STREAMS; ++x) want_att |= att_out[x] != nullptr;
  CFB_CHECK(!want_att || full_last, "cfb_sample: attention maps come from the full-cond branch; evaluate it (full_last=1)");
  cudaStream_t user_st = (cudaStream_t)stream;
  cudaStream_t st = user_st;
  const bool fenced = use_graph && (user_st == nullptr || user_st == cudaStreamLegacy || user_st == cudaStreamPerThread);
  if (fenced) {
    if (!h->own_stream) {
      CFB_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
      CFB_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
      CFB_CUDA(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
    }
    st = h->own_stream;
    CFB_CUDA(cudaEventRecord(h->ev_in, user_st));
    CFB_CUDA(cudaStreamWaitEvent(st, h->ev_in, 0));
  }
  const int S = sched->n_steps, n_batch = n_clips * n_branch;
  const int n_per_clip = h->ntok * h->lat;
  CFB_TRY(reserve_rows(h, n_batch, n_clips));
  CFB_TRY(h->tsteps.reserve((size_t)S * 4, &h->epoch));
  CFB_TRY(h->coef.reserve((size_t)(S + 1) * 8 * 4, &h->epoch));
  CFB_TRY(h->step.reserve(4, &h->epoch));
  CFB_TRY(h->x.reserve((size_t)n_clips * n_per_clip * 4, &h->epoch));
  // Schedule tables (timesteps -> time embedding / TimeBlock modulation tables, scheduler coefficients) are a pure
  // function of the schedule: they stay on the device across calls and are rebuilt only when the schedule changes
  // (a pageable-memory upload synchronises the stream, and prep_time is four launches).
  std::vector<float> tf(S);
  for (int i = 0; i < S; ++i) tf[i] = (float)sched->timesteps[i];
  const bool same_sched = h->sched_epoch == h->epoch && h->sched_ts == tf && (int)h->sched_coef.size() == S * 8 &&
                          memcmp(h->sched_coef.data(), sched->coef, (size_t)S * 8 * 4) == 0;
  if (!same_sched) {
    CFB_CUDA(cudaMemcpyAsync(h->tsteps.p, tf.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
    CFB_CUDA(cudaMemcpyAsync(h->coef.p, sched->coef, (size_t)S * 8 * 4, cudaMemcpyHostToDevice, st));
    CFB_CUDA(cudaMemsetAsync(h->coef.as<float>() + (size_t)S * 8, 0, 8 * 4, st));
    CFB_CUDA(cudaStreamSynchronize(st));   // tf is a host temporary
    CFB_TRY(prep_time(h, S, st));
    h->sched_ts = tf;
    h->sched_coef.assign(sched->coef, sched->coef + (size_t)S * 8);
    h->sched_epoch = h->epoch;             // prep_time may have grown buffers: remember the epoch AFTER it
    if (!h->ev_sched) CFB_CUDA(cudaEventCreateWithFlags(&h->ev_sched, cudaEventDisableTiming));
    CFB_CUDA(cudaEventRecord(h->ev_sched, st));
  } else {
    CFB_CUDA(cudaStreamWaitEvent(st, h->ev_sched, 0));   // a later call may arrive on another stream
  }
  CFB_CUDA(cudaMemsetAsync(h->step.p, 0, 4, st));
  MemLayout ml; CrossArgs ca;
  CFB_TRY(prep_memory(h, mem, n_batch, &ml, &ca, st));
  SharedPlan sp;
  CFB_TRY(make_shared_plan(h, mem, n_batch, &sp, st));
  CFB_TRY(build_rowblock_programs(h, n_batch, n_clips, sp, want_att, st));
  CFB_TRY(reserve_split(h, n_batch, mem, sp.on));
  {
    const char* e = getenv("CFB_CHAINS");
    int want = h->chains_override > 0 ? h->chains_override : e ? atoi(e) : 6;
    if (want < 1) want = 1;
    if (want > cfb_denoiser::MAX_CHAINS) want = cfb_denoiser::MAX_CHAINS;
    h->n_chains = want;
    if (!h->ev_fork) CFB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    // Chains are not equally long: the chain of the audio-only branch carries the 161-key per-pair attention.
    // CFB_PRIO_CHAINS (bit mask of chain indices) / CFB_PRIO_SIDE give those streams the highest launch priority.
    int prio_lo = 0, prio_hi = 0;
    CFB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const int prio_mask = getenv("CFB_PRIO_CHAINS") ? atoi(getenv("CFB_PRIO_CHAINS")) : 0;
    const int prio_side = getenv("CFB_PRIO_SIDE") ? atoi(getenv("CFB_PRIO_SIDE")) : 0;
    for (int c = 1; c < want; ++c) {
      if (!h->chain_st[c])
        CFB_CUDA(cudaStreamCreateWithPriority(&h->chain_st[c], cudaStreamNonBlocking, (prio_mask >> c) & 1 ? prio_hi : prio_lo));
      if (!h->ev_join[c]) CFB_CUDA(cudaEventCreateWithFlags(&h->ev_join[c], cudaEventDisableTiming));
    }
    const char* o = getenv("CFB_OVERLAP");
    if (want > 1 && !(o && atoi(o) == 0)) {
      for (int c = 0; c < want; ++c) {
        if (!h->chain_st2[c])
          CFB_CUDA(cudaStreamCreateWithPriority(&h->chain_st2[c], cudaStreamNonBlocking,
                                                (prio_side || ((prio_mask >> c) & 1)) ? prio_hi : prio_lo));
        if (!h->ev_a[c]) CFB_CUDA(cudaEventCreateWithFlags(&h->ev_a[c], cudaEventDisableTiming));
        if (!h->ev_b[c]) CFB_CUDA(cudaEventCreateWithFlags(&h->ev_b[c], cudaEventDisableTiming));
      }
      for (int i = 0; i < 2; ++i) {
        if (!h->pre_st[i]) CFB_CUDA(cudaStreamCreateWithFlags(&h->pre_st[i], cudaStreamNonBlocking));
        if (!h->ev_pre[i]) CFB_CUDA(cudaEventCreateWithFlags(&h->ev_pre[i], cudaEventDisableTiming));
      }
      if (!h->ev_mh) CFB_CUDA(cudaEventCreateWithFlags(&h->ev_mh, cudaEventDisableTiming));
    }
  }
  for (int x = 0; x < CFB_N_STREAMS; ++x) {
    ca.att_batch_stride[x] = (long long)h->L * h->ntok * mem->len[x];
    ca.att_step_stride[x] = (long long)n_clips * ca.att_batch_stride[x];
  }
  ca.att_first_batch = (n_branch - 1) * n_clips;
  ca.step_ptr = h->step.as<int>();
  const int n_inpaint = preseq ? preseq_len * h->lat : 0;
  if (preseq) {
    CFB_TRY(h->preseq.reserve((size_t)n_clips * n_inpaint * 4, &h->epoch));
    CFB_TRY(h->inp_noise.reserve((size_t)n_clips * n_inpaint * 4, &h->epoch));
  }
  auto init_state = [&]() -> int {   // step counter, latents (+ the first inpainting) as the first step expects them
    CFB_CUDA(cudaMemsetAsync(h->step.p, 0, 4, st));
    CFB_CUDA(cudaMemcpyAsync(h->x.p, latents, (size_t)n_clips * n_per_clip * 4, cudaMemcpyDeviceToDevice, st));
    if (preseq) {
      CFB_CUDA(cudaMemcpyAsync(h->preseq.p, preseq, (size_t)n_clips * n_inpaint * 4, cudaMemcpyDeviceToDevice, st));
      CFB_TRY(inpaint_first(h->x.as<float>(), h->preseq.as<float>(), h->inp_noise.as<float>(), h->coef.as<float>(),
                            n_clips, n_per_clip, n_inpaint, st));
    }
    return CFB_OK;
  };
  CFB_TRY(init_state());
  StepArgs sa{};
  sa.eps = h->eps.as<float>(); sa.x = h->x.as<float>(); sa.noise = step_noise; sa.coef = h->coef.as<float>();
  sa.step_ptr = h->step.as<int>(); sa.step_inc = h->step.as<int>(); sa.record = record;
  sa.preseq = preseq ? h->preseq.as<float>() : nullptr; sa.inp_noise = h->inp_noise.as<float>();
  sa.n_branch = n_branch; sa.full_last = full_last; sa.n_clips = n_clips; sa.n_per_clip = n_per_clip; sa.n_inpaint = n_inpaint;
  sa.n_steps = S; sa.kind = sched->kind; sa.clip_sample = sched->clip_sample; sa.guidance_scale = sched->guidance_scale;

  auto body = [&]() {
    return h->prec == CFB_BF16 ? step_body<bf16>(h, n_clips, n_branch, ml, ca, want_att ? att_out : nullptr, sa, sp, st)
                               : step_body<float>(h, n_clips, n_branch, ml, ca, want_att ? att_out : nullptr, sa, sp, st);
  };

  // fp32 on the tensor cores: the three-way splits of the weights are computed on first use and cached, which must not
  // happen inside a stream capture (the split would be replayed with every step) nor race between chain streams.  One
  // eager, single-stream evaluation of the step fills the cache before every (re)capture; the state it advanced is then
  // set up again.  Eager runs keep to one chain for the same reason.
  auto warm_splits = [&]() -> int {
    const int keep = h->n_chains;
    h->n_chains = 1;
    const int rc = body();
    h->n_chains = keep;
    CFB_TRY(rc);
    CFB_TRY(init_state());
    CFB_CUDA(cudaStreamSynchronize(st));
    return CFB_OK;
  };
  if (h->fp32_tc && !use_graph) h->n_chains = 1;

  if (use_graph) {
    cfb_denoiser::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.epoch = h->epoch; key.n_clips = n_clips; key.n_branch = n_branch; key.full_last = full_last; key.n_steps = S; key.kind = sched->kind;
    key.clip = sched->clip_sample; key.preseq_len = preseq ? preseq_len : 0; key.scale = sched->guidance_scale;
    for (int x = 0; x < CFB_N_STREAMS; ++x) {
      key.n_slots[x] = mem->n_slots[x]; key.len[x] = mem->len[x]; key.has_mask[x] = mem->mask[x] != nullptr;
      key.att[x] = want_att ? att_out[x] : nullptr;
    }
    key.noise = step_noise; key.record = record;
    key.plan[0] = sp.on + 2 * h->n_chains + 64 * h->rb_mask + 512 * (h->fp32_tc ? 1 + h->split_scheme : 0) + 4096 * h->act_sites + (h->act_f16 << 20);   // (g_bf16_act_f16_exclude is an env-only, process-constant setting) key.plan[1] = sp.on ? sp.n_groups : 0;
    for (int z = 0; sp.on && z < sp.n_groups; ++z) {
      key.plan[2 + 3 * z] = sp.g_stream[z]; key.plan[3 + 3 * z] = sp.g_row_start[z]; key.plan[4 + 3 * z] = sp.g_rows[z];
    }
    if (!h->graph_valid || memcmp(&key, &h->graph_key, sizeof(key)) != 0) {
      if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
      h->graph_valid = false;
      if (h->fp32_tc) CFB_TRY(warm_splits());
      cudaGraph_t graph = nullptr;
      CFB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      t_capturing = true;   // captured, not launched
      int rc = body();
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      t_capturing = false;
      if (rc != CFB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) { set_error("cudaStreamEndCapture: %s", cudaGetErrorString(ce)); return CFB_ERR_CUDA; }
      size_t n_nodes = 0;
      cudaGraphGetNodes(graph, nullptr, &n_nodes);
      ce = cudaGraphInstantiate(&h->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { set_error("cudaGraphInstantiate: %s", cudaGetErrorString(ce)); return CFB_ERR_CUDA; }
      h->graph_key = key; h->graph_valid = true;
      h->graph_nodes = n_nodes;
      static const bool verbose = getenv("CFB_VERBOSE") && atoi(getenv("CFB_VERBOSE"));
      if (verbose)
        fprintf(stderr, "[cfb] handle %p: captured a step graph (%zu nodes; %d clips x %d branches, %d chains, plan %d)\n",
                (void*)h, n_nodes, n_clips, n_branch, h->n_chains, (int)sp.on);
    }
    for (int i = 0; i < S; ++i) CFB_CUDA(cudaGraphLaunch(h->graph_exec, st));
    g_launches += (unsigned long long)S * h->graph_nodes;
  } else {
    for (int i = 0; i < S; ++i) CFB_TRY(body());
  }
  CFB_CUDA(cudaMemcpyAsync(latents, h->x.p, (size_t)n_clips * n_per_clip * 4, cudaMemcpyDeviceToDevice, st));
  if (fenced) {
    CFB_CUDA(cudaEventRecord(h->ev_out, st));
    CFB_CUDA(cudaStreamWaitEvent(user_st, h->ev_out, 0));
  }
  return CFB_OK;
}

}  // extern "C"
