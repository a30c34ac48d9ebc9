"""convofusion_b200 -- B200-native (sm_100a) implementation of ConvoFusion's sampling hot path.

Public surface (mirrors the reference's operator surface):
  Denoiser, ConvoFusionVae            drop-in modules (identical state_dict keys and call signatures)
  DDIMScheduler, DDPMScheduler        diffusers-0.14-compatible scheduler mirrors
  ConvoFusionSampler                  test_diffusion_forward / unbounded synthesis orchestration
  MotionWriter                        pred.npy / gt.npy / att_*.npy output layout of base.py:save_npy, asynchronous D2H
  SamplerPool                         several independent batches in flight on one GPU (one handle + stream per lane)
  slice_windows, window_text          host-side window bookkeeping of unbounded synthesis (process_samples / process_text)
All arithmetic runs in lib/libconvofusion_b200.so (hand-written CUDA, C ABI in include/convofusion_b200.h).
"""
from .modules import ConvoFusionVae, Denoiser
from .schedulers import DDIMScheduler, DDPMScheduler
from .conditioning import AudioConvEncoder, T5TextEncoder, TextAudioController, TextAudioMotionFuser
from .sampler import ConvoFusionSampler, default_denoiser, default_scheduler, default_vae
from .postprocess import keypoints3d
from .pool import SamplerPool
from .writer import MotionWriter
from .windows import slice_windows, window_spans, window_text


ACT_SITE_QKV, ACT_SITE_TIMEBLOCK, ACT_SITE_LINEAR1, ACT_SITE_LATENT_PROJ = 1, 2, 8, 16


def set_bf16_activation_sites(mask: int) -> None:
    """bf16 handles: the LayerNorm outputs feeding the GEMM sites in `mask` (ACT_SITE_*) are kept as two bf16 terms per
    value (hi + lo, two MMAs per K step).  Default 16 (latent_proj, whose output is eps itself: 0.017 instead of 0.028
    latent L2 from the fp32 reference after DDIM-50 for -0.8 % throughput); 0 = none.
    Process-wide (cfb_set_bf16_activation_sites)."""
    from . import _lib
    _lib.check(_lib.lib().cfb_set_bf16_activation_sites(int(mask)))


def set_bf16_activation_f16(enabled) -> None:
    """bf16 handles: which 16-bit activation operands are kept as fp16 (11 significant bits) instead of bf16 (8) and fed
    to fp16 x fp16 tensor-core products -- same bytes and instruction counts (1-3 % slower on power-capped GPUs).  True (default) = every group, False = none (bf16
    everywhere, the round-1 behaviour), or a bit mask: 1 LayerNorm outputs, 2 shared-slot probabilities / values and the
    per-pair attention output, 4 norm2 output / per-step keys, 8 self-attention q / k / v, 16 per-pair attention operands.
    Process-wide (cfb_set_bf16_activation_f16)."""
    from . import _lib
    mask = (31 if enabled else 0) if isinstance(enabled, bool) else int(enabled)
    _lib.check(_lib.lib().cfb_set_bf16_activation_f16(mask))


def set_vae_f16(enabled: bool) -> None:
    """16-bit `ConvoFusionVae` handles: fp16 (default) or bf16 weights and activations.  Read when a module packs its
    weights: call `vae.pack()` (or change it before the first call) for it to take effect.  Process-wide
    (cfb_set_vae_f16)."""
    from . import _lib
    _lib.check(_lib.lib().cfb_set_vae_f16(int(bool(enabled))))


def set_bf16_activation_terms(terms: int) -> None:
    """Shorthand for `set_bf16_activation_sites`: 2 = every site (27), 1 = none (0)."""
    from . import _lib
    _lib.check(_lib.lib().cfb_set_bf16_activation_terms(int(terms)))

__all__ = ["Denoiser", "ConvoFusionVae", "DDIMScheduler", "DDPMScheduler", "ConvoFusionSampler",
           "AudioConvEncoder", "T5TextEncoder", "TextAudioController", "TextAudioMotionFuser",
           "default_denoiser", "default_vae", "default_scheduler", "keypoints3d", "SamplerPool", "MotionWriter", "slice_windows", "window_spans", "window_text",
           "set_bf16_activation_terms", "set_bf16_activation_sites", "set_bf16_activation_f16", "set_vae_f16"]
