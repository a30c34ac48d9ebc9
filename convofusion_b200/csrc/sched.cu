// Fused 7-branch modality-guidance combine + scheduler step (+ latent inpainting for the next step).
//
// Arithmetic is written with explicit round-to-nearest intrinsics (no FMA contraction) in exactly
// the association order of the reference (convofusion.py:527-541) and of diffusers' step(), so the
// result matches an eager float32 evaluation bit for bit given the same coefficient table.
// Memory-bound: reads n_branch eps + x (+ noise), writes x (+ record): (n_branch + 2) * 4 B / element.
#include "common.cuh"
#include "kernels.cuh"

namespace cfb {

namespace {

__global__ void __launch_bounds__(256) guidance_sched_kernel(StepArgs a) {
  pdl_sync();
  const int total = a.n_clips * a.n_per_clip;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int step = a.step_ptr ? *a.step_ptr : 0;
  const float* cf = a.coef + (size_t)step * 8;
  const float s = a.guidance_scale;
  const float e0 = a.eps[i];
  float eps = e0;
  if (a.n_branch > 1) {
    // noise_pred_X = guidance_scale * 1 * (e_X - e_uncond); summed left to right over the branches that were
    // evaluated.  A single-modality branch whose conditioning equals the unconditional constant (monadic clips:
    // the speaker stream, dataset.py:185-199) has e_X == e_uncond, its term is an exact zero and the caller drops it.
    const int n_cond = a.n_branch - 1 - (a.full_last ? 1 : 0);
    float acc = 0.0f;
    if (n_cond > 0) acc = __fmul_rn(s, __fsub_rn(a.eps[(size_t)1 * total + i], e0));
    for (int g = 2; g <= n_cond; ++g) acc = __fadd_rn(acc, __fmul_rn(s, __fsub_rn(a.eps[(size_t)g * total + i], e0)));
    if (a.full_last)   // guidance_scale * 0 * (e_full - e_uncond)
      acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(s, 0.0f), __fsub_rn(a.eps[(size_t)(a.n_branch - 1) * total + i], e0)));
    eps = __fadd_rn(e0, acc);
  }
  const float x = a.x[i];
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(cf[0], eps)), cf[1]);
  if (a.clip_sample) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  float prev;
  if (a.kind == CFB_SCHED_DDIM) prev = __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], eps));
  else prev = __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], x));
  if (cf[4] != 0.0f && a.noise) prev = __fadd_rn(prev, __fmul_rn(cf[4], a.noise[(size_t)step * total + i]));
  if (a.record) a.record[(size_t)step * total + i] = prev;
  // unbounded_synthesis.py:70-76 for the NEXT step: overwrite the first tokens with the noised preseq
  const int within = i % a.n_per_clip;
  if (a.preseq && within < a.n_inpaint && step + 1 < a.n_steps) {
    const float* cn = cf + 8;
    const int j = (i / a.n_per_clip) * a.n_inpaint + within;
    prev = __fadd_rn(__fmul_rn(cn[5], a.preseq[j]), __fmul_rn(cn[6], a.inp_noise[j]));
  }
  a.x[i] = prev;
}

__global__ void inpaint_first_kernel(float* x, const float* preseq, float* inp_noise, const float* coef, int n_clips,
                                     int n_per_clip, int n_inpaint) {
  pdl_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_clips * n_inpaint) return;
  const int b = j / n_inpaint, within = j % n_inpaint;
  const size_t i = (size_t)b * n_per_clip + within;
  // noise = init_noise[:, :pl] (latents alias init_noise at step 0) ...
  const float v = __fadd_rn(__fmul_rn(coef[5], preseq[j]), __fmul_rn(coef[6], x[i]));
  x[i] = v;
  // ... and the in-place write also replaced init_noise[:, :pl], which later steps reuse as "noise".
  inp_noise[j] = v;
}

__global__ void step_inc_kernel(int* p) {
  pdl_sync(); *p += 1; }

}  // namespace

int guidance_sched_step(const StepArgs& a, cudaStream_t st) {
  const int total = a.n_clips * a.n_per_clip;
  if (total <= 0) return CFB_OK;
  CFB_CHECK(a.n_branch >= 1 && a.n_branch <= CFB_N_BRANCH, "guidance: n_branch must be 1..7");
  launch_k(guidance_sched_kernel, ceil_div(total, 256), 256, 0, st, a);
  CFB_LAUNCH_CHECK();
  if (a.step_inc) {
    launch_k(step_inc_kernel, 1, 1, 0, st, a.step_inc);
    CFB_LAUNCH_CHECK();
  }
  return CFB_OK;
}

int inpaint_first(float* x, const float* preseq, float* inp_noise, const float* coef, int n_clips, int n_per_clip,
                  int n_inpaint, cudaStream_t st) {
  const int total = n_clips * n_inpaint;
  if (total <= 0) return CFB_OK;
  launch_k(inpaint_first_kernel, ceil_div(total, 256), 256, 0, st, x, preseq, inp_noise, coef, n_clips, n_per_clip, n_inpaint);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

}  // namespace cfb
