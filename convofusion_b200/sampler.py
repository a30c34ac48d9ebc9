"""Sampling orchestration: the reference's test_diffusion_forward / _diffusion_reverse /
diffusion_reverse_forecast / process_samples call chain (convofusion.py:817-1065, 391-549;
unbounded_synthesis.py:28-187, 244-512) for synthetic, already-featurised inputs; word-excitation guidance (focus
tokens) runs on the as-written per-step loop with the gradient from the CUDA library (weg.py).

`ConvoFusionSampler` holds the same sub-module names as the reference LightningModule (`denoiser`, `vae`,
`text_audio_encoder`, `condition_fuser`), so `load_state_dict(ckpt["state_dict"])` accepts a reference
checkpoint (its T5 body is stripped on save, base.py:83-104).
"""
from __future__ import annotations

import os

import inspect
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor, nn

from .conditioning import (TextAudioController, TextAudioMotionFuser, expand_guidance_batch, guidance_branches,
                           guidance_memory, guidance_slots, guidance_slots_host)
from .modules import ConvoFusionVae, Denoiser
from .schedulers import DDIMScheduler, DDPMScheduler
from .weg import DEFAULT_WEG_PARAMETERS, FORECAST_WEG_PARAMETERS, weg_pre_step

MOTION_FPS = 25.0        # configs/config_cf_beatdnd.yaml:78-87
WINDOW_FRAMES = 128


class Ablation:
    """Attribute bag standing in for cfg.TRAIN.ABLATION (configs/config_cf_beatdnd.yaml:41-48)."""
    SKIP_CONNECT = True
    VAE_TYPE = "convofusion"
    PE_TYPE = "convofusion"
    DIFF_PE_TYPE = "convofusion"
    MLP_DIST = False
    CAUSAL_ATTN = False


def default_denoiser(precision: str = "bf16") -> Denoiser:
    """configs/modules/denoiser.yaml at config_cf_beatdnd."""
    return Denoiser(ablation=Ablation(), nfeats=189, condition="text+audio", latent_dim=[1, 128], ff_size=1024,
                    num_layers=9, num_heads=4, dropout=0.1, normalize_before=True, activation="gelu",
                    flip_sin_to_cos=True, return_intermediate_dec=False, position_embedding="sine", arch="trans_dec",
                    freq_shift=0, guidance_scale=7.5, guidance_uncondp=0.1, text_encoded_dim=512,
                    audio_encoded_dim=512, nclasses=10, precision=precision)


def default_vae(precision: str = "bf16") -> ConvoFusionVae:
    """configs/modules/motion_vae.yaml at config_cf_beatdnd."""
    return ConvoFusionVae(ablation=Ablation(), nfeats=189, latent_dim=[1, 128], ff_size=1024, num_layers=5, num_heads=2,
                          dropout=0.1, arch="encoder_decoder", normalize_before=True, activation="gelu",
                          position_embedding="sine", laplace_kernel_size=5, precision=precision)


def default_scheduler(kind: str = "ddim"):
    """configs/modules/scheduler.yaml params; DDIM is the 50-step override BASELINE.json names."""
    kw = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=True)
    return DDIMScheduler(**kw) if kind == "ddim" else DDPMScheduler(variance_type="fixed_small", **kw)


class ConvoFusionSampler(nn.Module):
    def __init__(self, denoiser: Optional[Denoiser] = None, vae: Optional[ConvoFusionVae] = None, scheduler=None,
                 noise_scheduler=None, guidance_scale: float = 7.5, num_inference_timesteps: int = 50,
                 eta: float = 0.0, precision: str = "bf16", vae_precision: Optional[str] = None):
        """vae_precision: precision of the VAE handle when it differs from the denoiser's (also env
        CONVOFUSION_B200_VAE_PRECISION).  The decode runs once per pass, so `precision="bf16", vae_precision="fp32"` buys
        the decoder's fp32 accuracy for the final joints at a few per cent of the pass."""
        super().__init__()
        vae_precision = vae_precision or os.environ.get("CONVOFUSION_B200_VAE_PRECISION") or precision
        self.denoiser = denoiser if denoiser is not None else default_denoiser(precision)
        self.vae = vae if vae is not None else default_vae(vae_precision)
        self.text_audio_encoder = TextAudioController(512)
        self.condition_fuser = TextAudioMotionFuser(512, self.denoiser.latent_dim)
        self.scheduler = scheduler if scheduler is not None else default_scheduler("ddim")
        self.noise_scheduler = noise_scheduler if noise_scheduler is not None else default_scheduler("ddpm")
        self.guidance_scale = guidance_scale
        self.num_inference_timesteps = num_inference_timesteps
        self.eta = eta
        self.weg_parameters = {k: (dict(v) if isinstance(v, dict) else v) for k, v in DEFAULT_WEG_PARAMETERS.items()}   # :64
        self.weg_forecast_parameters = {k: (dict(v) if isinstance(v, dict) else v) for k, v in FORECAST_WEG_PARAMETERS.items()}
        self.clf_guidance_drops = 6                              # convofusion.py:60
        self.do_classifier_free_guidance = guidance_scale > 1.0   # convofusion.py:131
        self.latent_dim = [1, self.denoiser.latent_dim]

    def set_precision(self, precision: str):
        self.denoiser.set_precision(precision)
        self.vae.set_precision(precision)
        return self

    # ------------------------------------------------------------------ conditioning
    def encode_conditions(self, clip: Dict[str, Tensor], uncond_text: Tensor, uncond_text_attn: Tensor):
        """convofusion.py:909-973, de-duplicated: (enc 5 x [1+B, M_x, 512], masks)."""
        return guidance_memory(self.text_audio_encoder, self.condition_fuser, clip, uncond_text, uncond_text_attn)

    # ------------------------------------------------------------------ reverse loops
    def _weg_denoiser(self) -> Denoiser:
        """fp32 evaluation of the denoiser for the WEG gradient: the module itself, or (bf16 samplers) a twin over the
        SAME parameter tensors with its own fp32 handle, packed on first use and re-packed when the parameters change."""
        if self.denoiser.precision == "fp32":
            return self.denoiser
        twin = self.__dict__.get("_weg_twin")
        if twin is None:
            import copy
            twin = copy.copy(self.denoiser)            # shares _parameters / _modules, owns no device handle yet
            twin.precision = "fp32"
            object.__setattr__(self, "_weg_twin", twin)   # not a registered sub-module: state_dict stays the reference's
        return twin

    def _diffusion_reverse(self, encoder_hidden_states, lengths=None, cond_masks=dict(), focus_indices=[],
                           init_latents: Optional[Tensor] = None, step_noise: Optional[Tensor] = None,
                           weg_log: Optional[list] = None):
        """The reference's as-written loop (convofusion.py:391-549): one Denoiser.forward on the 7*B batch and one
        scheduler.step per timestep, attention maps kept per step.  Same arithmetic as `sample()`, ~140 launches and
        two host syncs per step; kept for drop-in parity with callers that pass a pre-expanded 7*B batch, and the path
        word-excitation guidance runs on: with `focus_indices` every step is preceded by the latent update of
        convofusion.py:437-496 (weg.weg_pre_step, `self.weg_parameters`)."""
        if not self.do_classifier_free_guidance:
            # convofusion.py:398-401,527: guidance_scale <= 1 runs ONE conditional branch and no combine
            raise NotImplementedError("guidance_scale <= 1 (no classifier-free guidance) is not on the B200 hot path")
        dev = encoder_hidden_states[0].device
        mult = self.clf_guidance_drops + 1
        bsz = encoder_hidden_states[0].shape[0] // mult
        latents = init_latents if init_latents is not None else torch.randn(
            (bsz, 16, self.latent_dim[-1]), device=dev, dtype=torch.float)
        latents = latents * self.scheduler.init_noise_sigma
        self.scheduler.set_timesteps(self.num_inference_timesteps)
        timesteps = self.scheduler.timesteps.to(dev)
        extra = {}
        if "eta" in set(inspect.signature(self.scheduler.step).parameters.keys()):
            extra["eta"] = self.eta
        attention_matrices = dict()
        use_weg = len(focus_indices) > 0
        if use_weg:
            weg_den, scale_range = self._weg_denoiser(), self.weg_parameters["scale_range"]
        for i, t in enumerate(timesteps):
            if use_weg:
                latents, scale_range = weg_pre_step(weg_den, latents, i, t, encoder_hidden_states, cond_masks,
                                                    focus_indices, self.weg_parameters, scale_range, len(timesteps),
                                                    mult, weg_log)
            x = torch.cat([latents] * mult)
            noise_pred, att_mats = self.denoiser(sample=x, timestep=t, encoder_hidden_states=encoder_hidden_states,
                                                 lengths=None, mem_mask_dict=cond_masks)
            attention_matrices[t.item()] = [a.chunk(mult)[-1] for a in att_mats]
            noise_pred = self._combine(noise_pred)
            if step_noise is not None:
                extra["variance_noise"] = step_noise[i]
            latents = self.scheduler.step(noise_pred, t, latents, **extra).prev_sample
        return latents.permute(1, 0, 2), attention_matrices

    def _diffusion_reverse_forecast(self, encoder_hidden_states, lengths=None, preseq: Optional[Tensor] = None,
                                    cond_masks=dict(), focus_indices=[], init_noise: Optional[Tensor] = None,
                                    weg_log: Optional[list] = None):
        """unbounded_synthesis.py:28-187 as written (per-step Denoiser.forward / scheduler.step on the 7*B batch, latent
        inpainting of `preseq` [B, pl, 128] incl. the aliasing of `latents` and `init_noise` at step 0, :66-76), with the
        word-excitation update of :78-142 before every guided step when `focus_indices` is given
        (`self.weg_forecast_parameters` = the script's hard-coded values).  Returns (latents [16,B,128], the last
        step's full-cond attention maps).  Without focus tokens `sample(preseq=...)` runs the same loop fused."""
        if not self.do_classifier_free_guidance:
            raise NotImplementedError("guidance_scale <= 1 (no classifier-free guidance) is not on the B200 hot path")
        dev = encoder_hidden_states[0].device
        mult = self.clf_guidance_drops + 1
        bsz = encoder_hidden_states[0].shape[0] // mult
        init_noise = (init_noise.clone() if init_noise is not None else torch.randn(
            (bsz, 16, self.latent_dim[-1]), device=dev, dtype=torch.float)) * self.scheduler.init_noise_sigma
        self.scheduler.set_timesteps(self.num_inference_timesteps)
        timesteps = self.scheduler.timesteps.to(dev)
        extra = {}
        if "eta" in set(inspect.signature(self.scheduler.step).parameters.keys()):
            extra["eta"] = self.eta
        use_weg = len(focus_indices) > 0
        weg_den = self._weg_denoiser() if use_weg else None
        latents, att_mats = init_noise, None                              # :66 (alias)
        for i, t in enumerate(timesteps):
            if preseq is not None:
                pl = preseq.shape[1]
                noised = self.noise_scheduler.add_noise(preseq.clone(), init_noise.clone()[:, :pl, :], t)   # :73-75
                latents[:, :pl, :] = noised                               # :76 (also rewrites init_noise at step 0)
            if use_weg:
                wp = self.weg_forecast_parameters
                latents, _ = weg_pre_step(weg_den, latents, i, t, encoder_hidden_states, cond_masks, focus_indices, wp,
                                          wp["scale_range"], len(timesteps), mult, weg_log)
            x = torch.cat([latents] * mult)
            noise_pred, att = self.denoiser(sample=x, timestep=t, encoder_hidden_states=encoder_hidden_states,
                                            lengths=None, mem_mask_dict=cond_masks)
            att_mats = [a.chunk(mult)[-1] for a in att]
            latents = self.scheduler.step(self._combine(noise_pred), t, latents, **extra).prev_sample
        return latents.permute(1, 0, 2), att_mats

    def _combine(self, noise_pred: Tensor) -> Tensor:
        # convofusion.py:527-541 through the fused kernel with identity scheduler coefficients:
        # x0 = (0 - (-1)*eps)/1, prev = 1*x0 + 0*eps.
        import ctypes as C
        from . import _lib
        mult = self.clf_guidance_drops + 1
        eps = noise_pred.detach().to(torch.float32).contiguous()
        B = eps.shape[0] // mult
        out = torch.zeros(B, *eps.shape[1:], device=eps.device)
        coef = torch.tensor([-1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0], device=eps.device)
        with torch.cuda.device(eps.device):
            _lib.check(_lib.lib().cfb_guidance_sched_step(eps.data_ptr(), out.data_ptr(), 0, coef.data_ptr(), mult, B,
                                                          out.numel() // B, _lib.SCHED_DDIM, 0,
                                                          float(self.guidance_scale), _lib.stream_ptr()))
        return out

    @staticmethod
    def speaker_is_unconditional(clip: Dict[str, Tensor], uncond_text: Tensor, uncond_text_attn: Tensor) -> bool:
        """True when the speaker text of EVERY clip is the unconditional prompt (monadic BEAT clips: dataset.py:185-199
        feeds '-'*10 for the absent speaker), i.e. guidance branch 3 (speaker-only) repeats branch 0.  Callers that
        know (the data loader does) pass clip["spk_is_uncond"]; otherwise the features are compared on the device,
        which costs one read-back per call."""
        hint = clip.get("spk_is_uncond")
        if hint is not None:
            return bool(hint)
        t, a = clip["text_spk"], clip["text_spk_attn"]
        same = (t == uncond_text.to(t.device).unsqueeze(0)).all() & (a == uncond_text_attn.to(a.device).unsqueeze(0)).all()
        return bool(same)

    @torch.no_grad()
    def sample(self, enc: Sequence[Tensor], masks: Dict[str, Optional[Tensor]], n_clips: int,
               init_latents: Tensor, preseq: Optional[Tensor] = None, step_noise: Optional[Tensor] = None,
               record: bool = False, return_attention: bool = False, use_graph: bool = True,
               spk_is_uncond: bool = False):
        """Fused path: the whole reverse loop in one C call.  enc/masks are the de-duplicated slots from
        `encode_conditions`.  Returns (z [16,B,128] like _diffusion_reverse, record, attention).

        Branches evaluated (`guidance_branches`): the weight-0 full-cond branch only when its attention maps are asked
        for (convofusion.py:539); the speaker-only branch only when the speaker stream is conditional
        (`spk_is_uncond=False`).  Both omissions are exact: the dropped terms are zeros in convofusion.py:527-541."""
        if not self.do_classifier_free_guidance:
            # convofusion.py:398-401,527: guidance_scale <= 1 runs ONE conditional branch and no combine
            raise NotImplementedError("guidance_scale <= 1 (no classifier-free guidance) is not on the B200 hot path")
        branches = guidance_branches(return_attention, spk_is_uncond)
        slots = guidance_slots(n_clips, branches, init_latents.device)
        x = init_latents * self.scheduler.init_noise_sigma
        lat, rec, att = self.denoiser.sample(self.scheduler, enc, masks, slots, x, self.num_inference_timesteps,
                                             guidance_scale=self.guidance_scale, eta=self.eta, n_branch=len(branches),
                                             full_last=return_attention, slots_host=guidance_slots_host(n_clips, branches),
                                             step_noise=step_noise, preseq=preseq, noise_scheduler=self.noise_scheduler,
                                             record=record, return_attention=return_attention, use_graph=use_graph)
        return lat.permute(1, 0, 2), rec, att

    # ------------------------------------------------------------------ decode
    @torch.no_grad()
    def decode(self, z: Tensor, lengths: List[int]) -> Tensor:
        """convofusion.py:1027-1032: z [16,B,128] -> joints [B,T,189]."""
        ntok, bs, dim = z.shape
        zz = z.reshape(ntok // 2, 2, bs, dim).permute(1, 2, 0, 3)
        return self.vae.decode(zz, lengths)

    @torch.no_grad()
    def generate(self, clip: Dict[str, Tensor], uncond_text: Tensor, uncond_text_attn: Tensor, lengths: List[int],
                 init_latents: Tensor, **kw):
        """test_diffusion_forward (convofusion.py:817-1036) for one batch of featurised clips."""
        focus_indices = kw.pop("focus_indices", [])
        if len(focus_indices) > 0 and len(focus_indices[0]) > 0:      # convofusion.py:942-1023: WEG on the as-written loop
            enc, masks = self.encode_conditions(clip, uncond_text, uncond_text_attn)
            enc7, masks7 = expand_guidance_batch(enc, masks, init_latents.shape[0])
            z, att = self._diffusion_reverse(enc7, lengths, masks7, focus_indices=focus_indices, init_latents=init_latents,
                                             weg_log=kw.pop("weg_log", None))
            return {"m_rst": self.decode(z, lengths), "lat_t": z, "record": None, "test_attention_maps": att}
        kw.setdefault("spk_is_uncond", self.speaker_is_unconditional(clip, uncond_text, uncond_text_attn))
        enc, masks = self.encode_conditions(clip, uncond_text, uncond_text_attn)
        z, rec, att = self.sample(enc, masks, init_latents.shape[0], init_latents, **kw)
        return {"m_rst": self.decode(z, lengths), "lat_t": z, "record": rec, "test_attention_maps": att}

    # ------------------------------------------------------------------ unbounded synthesis
    @torch.no_grad()
    def synthesize_unbounded(self, windows: Sequence[Dict[str, Tensor]], uncond_text: Tensor, uncond_text_attn: Tensor,
                             init_noise: Sequence[Tensor], use_graph: bool = True,
                             focus_indices: Optional[Sequence[Sequence[Sequence[int]]]] = None):
        """process_samples (unbounded_synthesis.py:244-512): serial windows at 50 % overlap; the last 8 latent tokens
        of window k are inpainted into the first 8 of window k+1 (:442-444, 70-76) and the root x/z of every decoded
        window is re-anchored on the previous one (:461-465).  `windows[k]` is the featurised conditioning of
        window k for all B streams; returns the list of per-window joints [B,128,189].  `focus_indices[k]` (the word-
        excitation focus tokens process_samples derives for window k, :402-410; B = 1 like the reference) switches that
        window to the as-written loop with the WEG update (`_diffusion_reverse_forecast`)."""
        preseq, prev, outs = None, None, []
        for k, clip in enumerate(windows):
            B = clip["mel_lsn"].shape[0]
            enc, masks = self.encode_conditions(clip, uncond_text, uncond_text_attn)
            focus = focus_indices[k] if focus_indices is not None else []
            if len(focus) > 0 and len(focus[0]) > 0:
                enc7, masks7 = expand_guidance_batch(enc, masks, B)
                z, _ = self._diffusion_reverse_forecast(enc7, [WINDOW_FRAMES] * B, preseq, masks7, focus_indices=focus,
                                                        init_noise=init_noise[k])
            else:
                z, _, _ = self.sample(enc, masks, B, init_noise[k], preseq=preseq, use_graph=use_graph,
                                      spk_is_uncond=self.speaker_is_unconditional(clip, uncond_text, uncond_text_attn))
            preseq = z[z.shape[0] // 2:].permute(1, 0, 2).contiguous()
            feats = self.decode(z, [WINDOW_FRAMES] * B)
            if prev is not None:
                xz = torch.tensor([1.0, 0.0, 1.0], device=feats.device)
                feats[:, :, :3] = feats[:, :, :3] - feats[:, :1, :3] * xz
                feats[:, :, :3] = feats[:, :, :3] + prev[:, :1, :3] * xz
            outs.append(feats)
            prev = feats[:, WINDOW_FRAMES // 2:, :]
        return outs
