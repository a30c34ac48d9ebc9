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


def set_bf16_activation_terms(terms: int) -> None:
    """bf16 handles: 2 keeps the LayerNorm outputs as two bf16 terms per value (hi + lo) for the six GEMMs per layer they
    feed -- a third of the plain bf16 mode's deviation from the fp32 reference (0.068 vs 0.195 latent L2 after DDIM-50)
    at 0.8x its throughput; 1 (default) = plain bf16 operands.  Process-wide (cfb_set_bf16_activation_terms)."""
    from . import _lib
    _lib.check(_lib.lib().cfb_set_bf16_activation_terms(int(terms)))

__all__ = ["Denoiser", "ConvoFusionVae", "DDIMScheduler", "DDPMScheduler", "ConvoFusionSampler",
           "AudioConvEncoder", "T5TextEncoder", "TextAudioController", "TextAudioMotionFuser",
           "default_denoiser", "default_vae", "default_scheduler", "keypoints3d", "SamplerPool", "MotionWriter", "slice_windows", "window_spans", "window_text",
           "set_bf16_activation_terms"]
