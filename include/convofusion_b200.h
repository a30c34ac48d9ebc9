/*
 * convofusion_b200 -- C ABI of the B200-native ConvoFusion sampling hot path.
 *
 * The reference (m-hamza-mughal/convofusion) has no FFI: its extension point is the
 * `target:` string registry (convofusion/config.py:16-31) plus duck-typed calls on the
 * instantiated Python objects.  This header is therefore the boundary a maintainer
 * binds with ctypes; every entry point cites the reference
 * call it replaces.  Plain pointers and sizes only: all pointers are DEVICE pointers
 * unless a parameter is documented "host".  All functions return 0 on success, a
 * negative cfb_status otherwise; cfb_last_error() gives the message (thread local).
 *
 * Ownership: the caller owns every buffer it passes in (weights must stay alive and
 * unchanged for the lifetime of the handle that was created from them); handles own
 * their workspace, TMA descriptors and CUDA graphs.  One handle per (device, stream
 * of use); handles are not thread-safe.
 */
#ifndef CONVOFUSION_B200_H
#define CONVOFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFB_ABI_VERSION 2
#define CFB_N_STREAMS 5      /* spkemb, alsn, tlsn, apb, lsnemb: cross_attention.py:579 */
#define CFB_N_BRANCH 7       /* clf_guidance_drops + 1: convofusion.py:60,399-401 */

typedef enum {
  CFB_OK = 0,
  CFB_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  CFB_ERR_CUDA = -2,         /* CUDA runtime or driver error */
  CFB_ERR_NO_DEVICE = -3,    /* no sm_100 device: there is no CPU fallback */
  CFB_ERR_INTERNAL = -4
} cfb_status;

typedef enum { CFB_F32 = 0, CFB_BF16 = 1 } cfb_precision;
typedef enum { CFB_GEMM_AUTO = 0, CFB_GEMM_SIMT = 1, CFB_GEMM_TCGEN05 = 2 } cfb_gemm_backend;
typedef enum { CFB_ACT_NONE = 0, CFB_ACT_GELU = 1, CFB_ACT_SILU = 2, CFB_ACT_RELU = 3,
               CFB_ACT_LEAKY01 = 4 } cfb_act;
typedef enum { CFB_SCHED_DDIM = 0, CFB_SCHED_DDPM = 1 } cfb_sched_kind;

typedef void* cfb_stream;    /* cudaStream_t */

int cfb_abi_version(void);
const char* cfb_last_error(void);
/* Which GEMM engine bf16 contractions use (AUTO = tcgen05 whenever the shape allows). */
int cfb_set_gemm_backend(int backend);
/* Execution strategy of cfb_sample when slot tables are given: 1 (default) = shared-slot plan (memory-side
 * pre-projection of the unconditional slot), 0 = general per-pair path.  Same results up to
 * floating-point summation order; exposed so the tests can compare the two. */
int cfb_set_shared_plan(int enabled);
/* Row-block kernel (one persistent tcgen05 CTA per 128 query rows runs a layer's residual chains -- GEMM, residual
 * add, LayerNorm / TimeBlock modulation / SiLU, next GEMM -- on rows resident in tensor memory) in cfb_sample's bf16
 * step: bit 0 out_proj -> time_block1 -> norm2, bit 1 cross-attention values / fuser -> time_block2 -> norm3,
 * bit 2 linear2 -> next norm1.  Default 0 = one kernel per operator (the row-block programs are verified but slower at
 * the reference's batch sizes); 7 = all.  Same arithmetic up to summation order and the LayerNorm
 * statistics formula. */
int cfb_set_rowblock(int mask);
/* fp32 handles (precision = CFB_F32): 1 = the denoiser's GEMMs run on the tcgen05 tensor cores as three-way bf16 splits
 * of both fp32 operands (hi + mid + lo, six partial products, fp32 accumulation in tensor memory: fp32-level error at
 * 6x the MMA work of the bf16 mode, csrc/gemm_split.cu) instead of the CUDA-core FFMA GEMM (the default; env
 * CFB_FP32_TC); 0 = CUDA cores.  Modes 2 and 3 are precision-study settings, not production modes: 2 = activations
 * hi + lo (two bf16 terms) against bf16-rounded weights, 3 = both operands rounded to bf16 (the bf16 mode's GEMM
 * rounding with everything else in fp32).  Same operator surface as Denoiser.forward in float32 (denoiser.py:173-386). */
int cfb_set_fp32_tensor_cores(int mode);
/* bf16 handles, precision of the 16-bit ACTIVATION operands.  Activation rounding is what the guidance weights
 * (-36.5 / +7.5) amplify; weight rounding is common to all branches and cancels.
 *
 * cfb_set_bf16_activation_f16(mask): activation operands kept as fp16 (11 significant bits) instead of bf16 (8) --
 * they are O(1) by construction (LayerNorm outputs, probabilities, convex combinations of normalised memory, q / k / v;
 * every store clamps to the fp16 range) -- and consumed by fp16 x fp16 products: tcgen05.mma kind::f16 with both
 * operand formats f16 against fp16 copies of the bf16 weights made once per handle (exact above the fp16 subnormal
 * range; kind::f16 cannot mix an f16 A with a bf16 B), or the f16 form of the mma.sync attention kernels.  Same bytes,
 * same instruction counts, same kernels.  Bit mask of operand groups (default 31 = all, also env CFB_BF16_ACT_F16):
 *    1  LayerNorm outputs feeding qkv, both TimeBlock linears, linear1, latent_proj
 *    2  shared-slot probabilities with the per-step pre-projected values; per-pair attention output feeding the fuser
 *    4  norm2 output with the per-step pre-projected keys (scores) and the conditional-query projection
 *    8  q / k / v of the self-attention
 *   16  queries and normalised memory of the per-pair attention (and the memory-side pre-projections that read it)
 * DDIM-50 deviation from the fp32 reference (latent L2): 0.195 (mask 0) -> 0.084 (1) -> 0.049 (3) -> 0.028 (31) at
 * the same instruction counts (-0.5 to -3 % throughput on power-capped boxes: the fp16 products draw more).  0 = bf16 activations everywhere (round-1 behaviour).  Groups 2 (attention output) and 16 need
 * the mma.sync per-pair kernel: they are off while cfb_set_cross_tc(1); everything is off with cfb_set_rowblock != 0
 * and on the CUDA-core GEMM backend.
 *
 * cfb_set_bf16_activation_sites(mask): consumer sites (1 qkv, 2 both TimeBlock linears, 8 linear1, 16 latent_proj)
 * whose LayerNorm input is kept as TWO bf16 terms per value (hi + lo, 16 significant bits); that GEMM issues two
 * accumulating tcgen05.mma per K step against the bf16 weights.  Default 16 (also env CFB_BF16_ACT_SITES):
 * latent_proj's output IS eps, so its operand rounding reaches the guidance combine unattenuated -- 0.028 -> 0.017
 * for -0.8 % throughput.  With fp16 off: 16 -> 0.117, 18 -> 0.076 (-8 %), 27 -> 0.078 (-17 %).
 *
 * cfb_set_bf16_activation_terms: shorthand, 2 = every site (mask 27), 1 = none (mask 0). */
int cfb_set_bf16_activation_f16(int mask);
int cfb_set_bf16_activation_sites(int mask);
int cfb_set_bf16_activation_terms(int terms);
/* Per-pair cross-attention of bf16 handles (cross_attention.py:593-626: 16 queries against one clip's <= 256 memory
 * tokens): 1 = the tcgen05 / tensor-memory / TMA kernel (csrc/cross_tc.cu: both products issued transposed, queries as
 * the N = 16 dimension), 0 (default, also env CFB_CROSS_TC) = the mma.sync m16n8k16 kernel, which is faster for these
 * 16-row problems. */
int cfb_set_cross_tc(int enabled);
/* Kernels launched by this library since process start (bench.py's gpu_launches). */
unsigned long long cfb_launch_count(void);

/* ------------------------------------------------------------------ denoiser ----- */
/* Replaces Denoiser.forward (convofusion/models/architectures/denoiser.py:173-386) and the
 * 9 x TransformerDecoderLayer2Att.forward_pre it drives (operator/cross_attention.py:556-664).
 * Weights arrive PACKED by the host (convofusion_b200/pack.py): matrices are row-major
 * [out, in] in the handle's precision (float or bf16); biases / LayerNorm params float.
 * The five cross-attentions are algebraically folded (DESIGN.md "Folded cross-attention"):
 *   w_qx  [5*d, d]  = stack_x (W_k,x diag(gamma_x))^T W_q,x / sqrt(d)
 *   w_fu  [d, 5*d]  = cat_x  F_x O_x W_v,x diag(gamma_x)
 */
typedef struct {
  const float *ln1_g, *ln1_b;                 /* norm1 */
  const void  *w_in;  const float *b_in;      /* self_attn.in_proj   [3d, d] */
  const void  *w_so;  const float *b_so;      /* self_attn.out_proj  [d, d]  */
  const float *tb1_g, *tb1_b;                 /* time_block1.norm */
  const void  *w_tb1; const float *b_tb1;     /* time_block1.out_layers.2 [d, d] */
  const float *ln2_g, *ln2_b;                 /* norm2 */
  const void  *w_qx;  const float *b_qx;      /* folded query/key projection [5d, d] */
  const void  *w_fu;  const float *b_fu;      /* folded value/out/fuser      [d, 5d] */
  const float *tb2_g, *tb2_b;
  const void  *w_tb2; const float *b_tb2;
  const float *ln3_g, *ln3_b;                 /* norm3 */
  const void  *w_ff1; const float *b_ff1;     /* linear1 [ff, d] */
  const void  *w_ff2; const float *b_ff2;     /* linear2 [d, ff] */
} cfb_denoiser_layer;

typedef struct {
  int32_t d_model, latent_dim, n_tokens, n_layers, n_heads, ff_size, precision, pe_len;
  const void  *w_embed;                       /* latent_embd.weight [d, latent] */
  const float *tok_bias;                      /* [n_tokens, d] = latent_embd.bias + bh_embedding[tok%2] + query_pos.pe[tok/2] */
  const float *w_t1, *b_t1, *w_t2, *b_t2;     /* time_embedding.linear_{1,2} (always float) */
  const float *w_tbmod, *b_tbmod;             /* [n_layers*2*2d, d]: time_block{1,2}.emb_layers.1 stacked */
  const float *stream_emb;                    /* condition_embedding.weight [5, d] */
  const float *pe_mem;                        /* mem_pos.pe [pe_len, d] */
  const float *lnf_g, *lnf_b;                 /* decoder.norm */
  const void  *w_out; const float *b_out;     /* latent_proj [latent, d] */
  /* Shared-memory fast path (slot 0 of a stream attended by many batch entries, DESIGN.md "Shared slot"):
   * per stream x, stacked over layers l:  w_zx[x] [L*d, d] rows l*d.. = A_{x,l}^T  (A = w_qx block x),
   * a_zx[x] [L, d] = a_{x,l} (b_qx block x),  w_yx[x] [L*d, d] rows l*d.. = G_{x,l} (w_fu column block x). */
  const void  *w_zx[CFB_N_STREAMS];
  const float *a_zx[CFB_N_STREAMS];
  const void  *w_yx[CFB_N_STREAMS];
  const cfb_denoiser_layer *layers;           /* host array [n_layers] */
} cfb_denoiser_weights;

/* Conditioning memory of one call.  cond[x]: [n_slots[x], len[x], d] float, batch-first as
 * returned by TextAudioMotionFuser.forward (condfuser.py:32-51); mask[x]: [n_slots[x], len[x]]
 * bytes, 1 = ignore (key_padding_mask, cross_attention.py:587-591) or NULL; slot[x]: [n_rows/
 * n_tokens] int32 mapping each denoiser batch entry to a slot, or NULL for identity; slot_host[x]:
 * optional HOST copy of slot[x] (cfb_sample derives its execution plan from the tables: with a host
 * copy it does not have to read them back from the device). */
typedef struct {
  const float   *cond[CFB_N_STREAMS];
  const uint8_t *mask[CFB_N_STREAMS];
  const int32_t *slot[CFB_N_STREAMS];
  const int32_t *slot_host[CFB_N_STREAMS];
  int32_t n_slots[CFB_N_STREAMS];
  int32_t len[CFB_N_STREAMS];
} cfb_memory;

typedef struct cfb_denoiser cfb_denoiser;

int cfb_denoiser_create(const cfb_denoiser_weights *w, cfb_denoiser **out);
/* 16-bit handles: a second copy of the weight matrices in fp16, packed by the host FROM THE FP32 state_dict (same
 * struct, `precision` ignored, only the matrix pointers are read: layers[].w_in / w_so / w_tb1 / w_tb2 / w_qx / w_fu /
 * w_ff1 / w_ff2, w_embed, w_out, w_zx[], w_yx[]; the tensors must outlive the handle).  The fp16 x fp16 products of
 * cfb_set_bf16_activation_f16 then meet weights with 11 significant bits; without this call the handle uses fp16
 * conversions of its bf16 weights (8 bits).  Invalidates the handle's captured graph. */
int cfb_denoiser_attach_f16_weights(cfb_denoiser *h, const cfb_denoiser_weights *w16);
void cfb_denoiser_destroy(cfb_denoiser *h);

/* Concurrent chains of one captured sampling step: the guidance batch is cut into n_chains independent
 * row groups that run the layer stack side by side (1..8; 0 = library default: CFB_CHAINS or 6).
 * No reference counterpart (execution strategy only; results do not depend on it). */
int cfb_denoiser_set_chains(cfb_denoiser *h, int n_chains);

/* One denoiser evaluation: sample [n_batch, n_tokens, latent] float -> eps (same shape).
 * att_out[x] (or NULL): [n_batch, n_layers, n_tokens, len[x]] float attention weights, the
 * second return value of Denoiser.forward (cross_attention.py:234). */
int cfb_denoiser_forward(cfb_denoiser *h, const float *sample, int n_batch, int64_t timestep,
                         const cfb_memory *mem, float *eps_out, float *const att_out[CFB_N_STREAMS],
                         cfb_stream stream);

/* ------------------------------------------------------------------ sampling ----- */
/* Host-side schedule: one row of coefficients per inference step, computed by the host
 * scheduler mirror exactly like diffusers' float32 table arithmetic:
 *   coef[i] = { sqrt(1-abar_t), sqrt(abar_t), k0, k1, k2, ia, ib, 0 }
 *   DDIM: prev = k0*x0 + k1*eps (+ k2*noise if k2 != 0)   DDPM: prev = k0*x0 + k1*x (+ k2*noise)
 *   x0 = (x - coef0*eps)/coef1, clamped to [-1,1] when clip_sample
 *   ia, ib: add_noise coefficients of noise_scheduler at t_i (latent inpainting).
 */
typedef struct {
  int32_t kind;              /* cfb_sched_kind */
  int32_t n_steps;
  int32_t clip_sample;
  float   guidance_scale;
  const int64_t *timesteps;  /* host [n_steps] */
  const float   *coef;       /* host [n_steps, 8] */
} cfb_schedule;

/* Word-excitation guidance (Convofusion._diffusion_reverse with focus tokens, convofusion.py:437-496; the iterative
 * refinement :298-388; tools/word_excitation_guidance.py:55-62 update_latent = torch.autograd.grad through
 * Denoiser.forward).  fp32 handles.  _forward evaluates the denoiser on `sample` [n_batch, n_tokens, latent] (the
 * text-only branch: n_batch = clips), keeps the activations and writes the attention maps of stream `att_stream`
 * (2 = listener text) as [n_batch, n_layers, n_tokens, len] -- the tensor the reference's loss is computed from.
 * _backward takes dLoss/dAtt in the same layout and returns dLoss/dsample [n_batch, n_tokens, latent]. */
int cfb_denoiser_weg_forward(cfb_denoiser *h, const float *sample, int n_batch, int64_t timestep, const cfb_memory *mem,
                             int att_stream, float *att_out, cfb_stream stream);
int cfb_denoiser_weg_backward(cfb_denoiser *h, const float *d_att, float *grad_sample, cfb_stream stream);

/* Replaces Convofusion._diffusion_reverse (modeltype/convofusion.py:391-549) and
 * diffusion_reverse_forecast (unbounded_synthesis.py:28-187) with WEG off: the whole
 * n_steps loop (7-branch guidance + scheduler step [+ latent inpainting]) on the device.
 *   n_clips      B; the denoiser batch is n_branch*B rows of n_tokens, branch-major like
 *                torch.cat([latents]*7) (convofusion.py:499)
 *   n_branch     guidance branches evaluated (1..7), in the reference's order (convofusion.py:910) with any subset of
 *                the single-modality branches left out: a branch whose conditioning equals the unconditional
 *                constant contributes guidance_scale * (e_uncond - e_uncond) = 0 exactly (monadic BEAT clips: the
 *                speaker branch, dataset.py:185-199).  Branch 0 is always the all-unconditional one.
 *   full_last    1 = the last evaluated branch is the weight-0 full-cond branch (convofusion.py:539; needed only for
 *                attention maps), 0 = it is skipped
 *   mem          slot[x] has n_branch*B entries
 *   latents      in: initial noise * init_noise_sigma [B, n_tokens, latent]; out: final latents
 *   step_noise   [n_steps, B, n_tokens, latent] or NULL (DDPM / eta>0 noise)
 *   preseq       [B, preseq_len, latent] or NULL (previous window's latents to inpaint)
 *   record       [n_steps, B, n_tokens, latent] or NULL: latents after every scheduler step
 *   att_out[x]   [n_steps, B, n_layers, n_tokens, len[x]] or NULL: maps of the LAST branch
 *   use_graph    replay one captured CUDA graph per step instead of launching kernels
 */
int cfb_sample(cfb_denoiser *h, const cfb_schedule *sched, const cfb_memory *mem, int n_clips,
               int n_branch, int full_last, float *latents, const float *step_noise, const float *preseq,
               int preseq_len, float *record, float *const att_out[CFB_N_STREAMS], int use_graph,
               cfb_stream stream);

/* Fused 7-way guidance combine + scheduler step (convofusion.py:527-545 + diffusers step()).
 * eps [n_branch, B, n] ; x [B, n] in/out; coef = one device row of 8 floats as above.
 * n_branch 1..7: branch 0 unconditional, the others single-modality branches; 7 = the last one is the weight-0
 * full-cond branch. */
int cfb_guidance_sched_step(const float *eps, float *x, const float *noise, const float *coef_dev,
                            int n_branch, int n_clips, int n_per_clip, int kind, int clip_sample,
                            float guidance_scale, cfb_stream stream);

/* ------------------------------------------------------------------ VAE decode --- */
/* Replaces ConvoFusionVae.decode (architectures/vae.py:268-372): SkipTransformerDecoder x2
 * (cross_attention.py:89-125, 361-382) + final linears + padding mask. */
typedef struct {
  const float *ln1_g, *ln1_b; const void *w_in; const float *b_in;   /* self_attn.in_proj [3d,d] */
  const void  *w_so; const float *b_so;
  const float *ln2_g, *ln2_b; const void *w_q;  const float *b_q;    /* multihead_attn q rows [d,d] */
  const void  *w_kv; const float *b_kv;                              /* multihead_attn k,v rows [2d,d] */
  const void  *w_co; const float *b_co;
  const float *ln3_g, *ln3_b; const void *w_ff1; const float *b_ff1;
  const void  *w_ff2; const float *b_ff2;
} cfb_vae_layer;

typedef struct {
  const cfb_vae_layer *layers;     /* host [n_layers]: input_blocks.., middle_block, output_blocks.. */
  const void  *w_skip[4]; const float *b_skip[4];   /* linear_blocks.i [d, 2d], (n_layers-1)/2 used */
  const float *lnf_g, *lnf_b;      /* decoder.norm */
  const void  *w_final; const float *b_final; int32_t n_out;   /* {body,hands}_final_layer [n_out, d] */
} cfb_vae_decoder;

/* Encode side (architectures/vae.py:162-266): SkipTransformerEncoder of pre-norm TransformerEncoderLayers
 * (cross_attention.py:41-64, 288-300) over 2 distribution tokens + 16 frames per 16-frame chunk. */
typedef struct {
  const float *ln1_g, *ln1_b; const void *w_in; const float *b_in;   /* self_attn.in_proj [3d,d] */
  const void  *w_so; const float *b_so;
  const float *ln2_g, *ln2_b; const void *w_ff1; const float *b_ff1;
  const void  *w_ff2; const float *b_ff2;
} cfb_vae_enc_layer;

typedef struct {
  const cfb_vae_enc_layer *layers;  /* host [n_layers] or NULL when the encode side is not provided */
  const void  *w_skip[4]; const float *b_skip[4];
  const float *lnf_g, *lnf_b;       /* encoder.norm */
  const float *tokens;              /* {body,hands}_global_motion_token [2, d] */
  const float *w_emb, *b_emb;       /* {body,hands}_skel_embedding [d, n_in] (always float) */
  int32_t n_in, col0;               /* feature columns [col0, col0 + n_in) of the 189-wide input */
} cfb_vae_encoder;

typedef struct {
  int32_t d_model, n_layers, n_heads, ff_size, precision, pe_len;
  const float *pe_query, *pe_mem;   /* query_pos_decoder.pe / mem_pos_decoder.pe [pe_len, d] */
  cfb_vae_decoder part[2];          /* body, hands */
  const float *pe_enc;              /* query_pos_encoder.pe [pe_len, d] */
  cfb_vae_encoder enc[2];           /* body, hands */
} cfb_vae_weights;

typedef struct cfb_vae cfb_vae;
int cfb_vae_create(const cfb_vae_weights *w, cfb_vae **out);
/* 16-bit VAE handles (precision = CFB_BF16): 1 (default, also env CFB_VAE_F16) = weights and activations are fp16
 * instead of bf16 -- the host packs the matrices as fp16 (ask cfb_get_vae_f16 when packing), GEMMs run tcgen05.mma
 * kind::f16 on f16 operands, the attention kernels their f16 mma.sync form; 11 instead of 8 significant bits at the
 * same bytes and speed (decode max-relative error 6.6e-3 -> see DESIGN.md section 2).  The VAE has no guidance
 * amplification, so here the weight format matters as much as the activations'.  Read when a handle is created. */
int cfb_set_vae_f16(int enabled);
int cfb_get_vae_f16(void);
void cfb_vae_destroy(cfb_vae *h);
/* z [2, B, n_chunks, d] float; lengths host [B]; out [B, n_frames, n_out_body+n_out_hands]. */
int cfb_vae_decode(cfb_vae *h, const float *z, int n_clips, int n_chunks, int n_frames,
                   const int32_t *lengths_host, float *out, cfb_stream stream);
/* Replaces the deterministic part of ConvoFusionVae.encode (vae.py:162-258): features [B, T, 189] (T a multiple of
 * 16) -> mu, std [2, B*T/16, d] (body then hands; torch.distributions.Normal(mu, std) is built by the caller) and the
 * chunk-root-subtracted features [B, T, 189] (third return value of encode). */
int cfb_vae_encode(cfb_vae *h, const float *features, int n_clips, int n_frames, const int32_t *lengths_host,
                   float *mu_out, float *std_out, float *feats_out, cfb_stream stream);

/* ------------------------------------------------------------------ unit ops ----- */
/* Exposed so tests/ can check each kernel against the oracle through the same ABI.
 * y[M,N] (+)= act(A[M,K] W[N,K]^T + bias).  a_bf16/out_bf16 pick element types. */
int cfb_linear(const void *A, int a_bf16, const void *W, const float *bias, void *out, int out_bf16,
               int M, int N, int K, int act, int a_act, int accumulate, int backend, cfb_stream stream);
int cfb_layernorm(const float *x, const float *g, const float *b, void *out, int out_bf16,
                  int rows, int d, cfb_stream stream);
/* Generic multi-head attention over packed projections (torch.nn.MultiheadAttention core):
 * q [n*Lq, ldq], k/v [n*Lk, ldk] rows are sample-major; kv_len [n] or NULL. */
int cfb_mha(const void *q, int ldq, const void *k, const void *v, int ldk, void *out, int ldo,
            int is_bf16, int n, int Lq, int Lk, int n_heads, int head_dim, const int32_t *kv_len,
            cfb_stream stream);
/* Conditioning projections (audioenc.py:29-34; t5.py:48-49,57), float in/out. */
int cfb_audio_encoder(const float *mel, int rows, const float *w0, const float *b0, const float *w1,
                      const float *b1, const float *w2, const float *b2, int n_mel, int hidden,
                      int d_out, float *tmp0, float *tmp1, float *out, cfb_stream stream);

/* Output post-processing of the test / demo writers (models/modeltype/base.py:204-209): features [rows, 63*3]
 * -> keypoints [rows, 63, 3] = feats / 3, fingers re-attached to their wrist (joints 43.. += joint 11,
 * joints 23..42 += joint 7, both before the root is added), then every joint but the root += root. */
int cfb_keypoints3d(const float *feats, long long rows, float *out, cfb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* CONVOFUSION_B200_H */
