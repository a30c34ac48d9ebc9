// Internal launcher declarations (definitions in rowops.cu / attention.cu / sched.cu).
#pragma once
#include "common.cuh"

namespace cfb {

// ---- rowops.cu
template <typename T>
int ln_rows(const float* x, const float* g, const float* b, const float* mod, const int* step_ptr,
            long long mod_step_stride, T* out, int rows, int d, cudaStream_t st, int terms = 1);
// 16-bit payload conversion bf16 -> fp16 (values clamped to the fp16 range; exact above its subnormal range)
int bf16_to_f16(const bf16* in, bf16* out, size_t n, cudaStream_t st);
int mem_build(const float* const cond[CFB_N_STREAMS], const int n_slots[CFB_N_STREAMS], const int len[CFB_N_STREAMS],
              const float* stream_emb, const float* pe, float* mem_c, int d, cudaStream_t st);
template <typename T>
int mem_hat(const float* mem_c, const float* temb, const int* step_ptr, T* out, int rows, int d, cudaStream_t st,
            int f16 = 0);   // f16: 16-bit output holds fp16
int time_sinusoid(const float* t, float* out, int n, int dim, cudaStream_t st);
template <typename T> int cast_rows(const float* in, T* out, long long n, cudaStream_t st);
template <typename T> int concat2(const float* a, const float* b, T* out, int rows, int d, cudaStream_t st);
template <typename T> int add_pe(const float* src, const float* pe, T* out, int n_batch, int L, int d, cudaStream_t st);
int mask_frames(float* out, const int* lengths, int n_batch, int L, int d, cudaStream_t st);
int keypoints3d(const float* feats, long long rows, float* out, cudaStream_t st);
// VAE encode side (vae.py:176-260)
int chunk_root(const float* in, float* out, long long n_rows, int nf, int chunk, cudaStream_t st);
int enc_assemble(const float* emb, const float* tokens, const float* pe, float* h, int n, int n_tok, int chunk, int d,
                 cudaStream_t st);
int enc_dist(const float* y, float* mu, float* sd, int n, int L, int d, cudaStream_t st);

// ---- attention.cu
// Generic MHA core on projected q/k/v (rows sample-major): softmax(q k^T / sqrt(hd) + pad) v.
template <typename T>
int mha(const T* q, int ldq, const T* k, const T* v, int ldk, T* out, int ldo, int n, int Lq, int Lk, int n_heads,
        int head_dim, const int* kv_len, cudaStream_t st);
// bf16 handles with fp16 activations: q / k / v hold fp16, output bf16 or fp16 (attention.cu)
bool mha_f16_supported(int Lk, int head_dim);
int mha_f16(const bf16* q, int ldq, const bf16* k, const bf16* v, int ldk, bf16* out, int ldo, int n, int Lq, int Lk,
            int n_heads, int head_dim, cudaStream_t st, int out_f16 = 0);

// The denoiser's five folded single-head cross-attentions for every (batch entry, stream):
//   P = softmax(qx[bs, x] . mem_hat[slot]^T + mask),  u[bs, x] = P . mem_hat[slot]
struct CrossArgs {
  int row_base[CFB_N_STREAMS];          // first mem_hat row of stream x
  int len[CFB_N_STREAMS];
  const int* slot[CFB_N_STREAMS];       // [n_batch] or nullptr (identity)
  const uint8_t* mask[CFB_N_STREAMS];   // [n_slots, len] or nullptr
  float* att[CFB_N_STREAMS];            // attention-map output for this layer or nullptr
  long long att_batch_stride[CFB_N_STREAMS];  // elements between consecutive batch entries
  int att_first_batch;                  // maps are written for batch entries >= this
  long long att_step_stride[CFB_N_STREAMS];   // added per *step_ptr (graph replay)
  const int* step_ptr;
  int skip_slot0;                       // 1: (batch entry, stream) pairs on slot 0 are handled by the shared path
  int bs_offset;                        // first batch entry handled by this launch (blockIdx.x is relative to it)
  // tcgen05 per-pair kernel (cross_tc.cu): slots per stream and the transposed copy of the memory
  int n_slots[CFB_N_STREAMS];
  const bf16* mem_hat_t;                // per stream x at element offset t_off[x]: [n_slots, 512, lenp[x]] or nullptr
  long long t_off[CFB_N_STREAMS];
  int lenp[CFB_N_STREAMS];              // len rounded up to a multiple of 8 (16-byte rows)
  int out_f16;                          // bf16 path: u is written as fp16 (mma.sync kernel only; A operand of an fp16 GEMM)
  int in_f16;                           // bf16 path: qx and mem_hat hold fp16 (mma.sync kernel only: f16 MMAs)
};
// cross_tc.cu
int init_cross_tc_kernels();
bool cross_tc_supported(const CrossArgs& a, int n_tokens, int d);
int mem_transpose(const bf16* mem_hat, bf16* mem_hat_t, const CrossArgs& a, cudaStream_t st);
int cross_attention_tc(const bf16* qx, int q_rows, const bf16* mem_hat, bf16* u, const CrossArgs& a, int n_batch,
                       cudaStream_t st);
// Shared-slot path: scores of every row against slot 0 of every stream come from ONE GEMM (S, fp32, stream x
// at columns s_off[x] .. +len[x]); this turns them into bf16 probabilities P (stream x at p_off[x] .. +kp[x],
// zero padded), writing zeros where the pair is conditional (slot != 0) and handled by cross_attention().
struct SharedAttnArgs {
  int len[CFB_N_STREAMS], s_off[CFB_N_STREAMS], p_off[CFB_N_STREAMS], kp[CFB_N_STREAMS];
  const int* slot[CFB_N_STREAMS];       // [n_batch]
  const uint8_t* mask[CFB_N_STREAMS];   // key padding mask of slot 0 or nullptr
  int ld_s, ld_p;
  int bs_offset;                        // batch entry of row 0 of S / P as passed to the launch
  int p_f16;                            // 16-bit P holds fp16 instead of bf16 (A operand of an fp16 GEMM)
};
template <typename TP>
int softmax_shared(const float* S, TP* P, const SharedAttnArgs& a, int n_batch, int n_tokens, cudaStream_t st);
template <typename T>
int shared_key_bias(const T* mem_hat, float* z0, const float* const a_zx[CFB_N_STREAMS],
                    const int row_base[CFB_N_STREAMS], const int len[CFB_N_STREAMS], const int s_off[CFB_N_STREAMS],
                    int n_layers, int n_tot, cudaStream_t st, int f16 = 0);   // f16: 16-bit mem_hat holds fp16
template <typename T>
int cross_attention(const T* qx, const T* mem_hat, T* u, const CrossArgs& a, int n_batch, int n_tokens, int d,
                    cudaStream_t st);

// ---- weg.cu: input-gradient kernels of the denoiser's query side (word-excitation guidance)
int linear_bwd(const float* dY, int ldy, const float* W, int ldw, float* dX, int ldx, int M, int N, int K, int accumulate,
               cudaStream_t st);
int ln_bwd(const float* x, const float* gamma, const float* beta, const float* mod, const float* dy, float* g, int rows, int d,
           cudaStream_t st);
int gelu_bwd(const float* z, float* df, long long n, cudaStream_t st);
int mha_bwd(const float* qkv, const float* dO, float* dqkv, int n_batch, int L, int n_heads, int d, cudaStream_t st);
int cross_bwd(const float* dcat, const float* mem_hat, float* dqx, const CrossArgs& a, const float* d_att, int att_stream,
              int n_batch, int n_tokens, cudaStream_t st);

// ---- sched.cu
struct StepArgs {
  const float* eps;        // [n_branch, B, n]
  float* x;                // [B, n] latents, updated in place
  const float* noise;      // [n_steps?, B, n] or nullptr
  const float* coef;       // device [n_steps, 8]
  const int* step_ptr;     // device step counter (nullptr = row 0)
  int* step_inc;           // if set, incremented by one after the update (graph replay)
  float* record;           // [n_steps, B, n] or nullptr
  const float* preseq;     // [B, pl*lat] inpainting source or nullptr
  const float* inp_noise;  // [B, pl*lat]
  int n_branch, n_clips, n_per_clip, n_inpaint, n_steps, kind, clip_sample;
  int full_last;           // the last evaluated branch is the weight-0 full-cond branch (convofusion.py:539)
  float guidance_scale;
};
int guidance_sched_step(const StepArgs& a, cudaStream_t st);
// Step-0 inpainting incl. the reference's aliasing quirk (unbounded_synthesis.py:66-76).
int inpaint_first(float* x, const float* preseq, float* inp_noise, const float* coef, int n_clips, int n_per_clip,
                  int n_inpaint, cudaStream_t st);

}  // namespace cfb
