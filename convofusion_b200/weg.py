"""Word-excitation guidance (WEG): the latent update that precedes every guided step when focus tokens are given.

Mirrors `convofusion/models/tools/word_excitation_guidance.py` (aggregate_attentions :11-14,
get_max_attention_at_indices :16-52, update_latent :55-62, compute_attention_focus_loss :65-83) and the loop code that
drives it (`Convofusion._diffusion_reverse` convofusion.py:437-496, `iterative_refinement_step` :298-388;
`diffusion_reverse_forecast` unbounded_synthesis.py:82-142 is the same block).

Split of work: the denoiser evaluation on the text-only branch and the gradient of the loss with respect to the latents
run in the CUDA library (`Denoiser.weg_forward` / `weg_backward`, csrc/weg.cu -- the reference uses torch.autograd
through Denoiser.forward); the loss itself is a function of one [1, 9, 16, T] attention tensor (layer mean, softmax over
the tokens between BOS and EOS, 3x3 Gaussian smoothing, max over the motion tokens, hinge), evaluated with torch on the
device, whose autograd supplies dLoss/dAtt for that small tensor only.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

# configs/assets.yaml:18-24
DEFAULT_WEG_PARAMETERS = {"scale_factor": 1000, "scale_range": [1.0, 0.5], "max_iter_to_alter": 800,
                          "thresholds": {0: 0.05, 200: 0.4, 400: 0.6, 600: 0.8}, "max_refinement_steps": 300}
# unbounded_synthesis.py:82-87: diffusion_reverse_forecast hard-codes its own ("TODO: move to config")
FORECAST_WEG_PARAMETERS = {"scale_factor": 100, "scale_range": (1.0, 0.5), "max_iter_to_alter": 800,
                           "thresholds": {0: 0.05, 200: 0.4, 400: 0.6, 600: 0.8}, "max_refinement_steps": 300}
TEXT_STREAM = 2   # memory order (spkemb, alsn, tlsn, apb, lsnemb): cross_attention.py:579


def gaussian_kernel(kernel_size: int = 3, sigma: float = 0.5) -> Tensor:
    """operator/gaussian_smoothing.py:21-47 (dim = 2, one channel), including its exp(-((x - mean) / (2 sigma))^2)."""
    kernel = torch.ones(())
    mean = (kernel_size - 1) / 2
    for mgrid in torch.meshgrid([torch.arange(kernel_size, dtype=torch.float32)] * 2, indexing="ij"):
        kernel = kernel * (1 / (sigma * math.sqrt(2 * math.pi)) * torch.exp(-((mgrid - mean) / (2 * sigma)) ** 2))
    return (kernel / kernel.sum()).view(1, 1, kernel_size, kernel_size)


def aggregate_attentions(att_mats: Tensor) -> Tensor:
    return torch.mean(att_mats, dim=1)                                        # :11-14


def get_max_attention_at_indices(att_mat: Tensor, batch_idxs: Sequence[Sequence[int]], smooth_attentions: bool = False,
                                 normalize_eot: bool = False, eot_indices=()) -> List[List[Tensor]]:
    last_idx = -1
    if normalize_eot:
        assert len(eot_indices) > 0, "Need to provide eot indices for normalization"
        assert att_mat.shape[0] == 1, "EOS/BOS normalization only works for test batch size 1 currently"
        last_idx = int(eot_indices[0])
    a = torch.softmax(att_mat[:, :, 1:last_idx], dim=-1)                      # :28-30
    if smooth_attentions:
        a = F.conv2d(F.pad(a.unsqueeze(1), (1, 1, 1, 1), mode="reflect"), gaussian_kernel().to(a)).squeeze(1)   # :33-36
    return [[a[b, :, i - 1].max(dim=-1)[0] for i in idxs] for b, idxs in enumerate(batch_idxs)]   # :39-51


def compute_attention_focus_loss(max_attention_at_indices: List[List[Tensor]]):
    losses = []
    for sample in max_attention_at_indices:
        if len(sample) == 0:
            raise ValueError("every clip needs at least one focus token (the reference's empty-sample branch is CUDA-only)")
        losses.append(torch.mean(torch.stack([torch.max(torch.zeros_like(t), 1.0 - t) for t in sample]), dim=-1))
    losses = torch.stack(losses, dim=-1)
    return torch.mean(losses), losses                                         # :80-83


class WegEvaluation:
    """One forward of the text-only branch with the focus loss; `grad()` is torch.autograd.grad(loss, latents)."""

    def __init__(self, denoiser, latents: Tensor, t, enc_text, masks_text, focus_indices, eot_indices):
        self.denoiser, self.latents = denoiser, latents
        att = denoiser.weg_forward(latents, int(t), enc_text, masks_text, stream=TEXT_STREAM)
        self.att = att.detach().requires_grad_(True)
        with torch.enable_grad():
            agg = aggregate_attentions(self.att)
            self.max_att = get_max_attention_at_indices(agg, focus_indices, smooth_attentions=True, normalize_eot=True,
                                                        eot_indices=eot_indices)
            self.loss, _ = compute_attention_focus_loss(self.max_att)

    def grad(self) -> Tensor:
        d_att, = torch.autograd.grad(self.loss, [self.att], retain_graph=True)
        return self.denoiser.weg_backward(d_att)


def weg_pre_step(denoiser, latents: Tensor, i: int, t, enc: Sequence[Tensor], masks: Dict[str, Optional[Tensor]],
                 focus_indices, weg_parameters: dict, scale_range, n_steps: int, mult: int = 7,
                 log: Optional[list] = None):
    """convofusion.py:437-496.  `enc` / `masks` hold the 7*B guidance batch (the text-only branch is chunk 1, :449-450);
    `scale_range` is re-assigned to the linspace ARRAY on every step exactly like the reference (:442-444).
    Returns (latents, scale_range)."""
    scale_range = np.linspace(scale_range[0], scale_range[1], n_steps)
    enc_t = [e.chunk(mult)[1] for e in enc]
    masks_t = {k: (v.chunk(mult)[1] if v is not None else v) for k, v in masks.items()}
    eot = (torch.argmax(masks_t["tlsn"].int(), dim=1) - 1).tolist()           # :463 (one host read per step)
    ev = WegEvaluation(denoiser, latents, t, enc_t, masks_t, focus_indices, eot)
    thresholds = weg_parameters["thresholds"]
    n_refine = 0
    if i in thresholds.keys() and float(ev.loss) > 1.0 - thresholds[i]:       # :474
        step_size = weg_parameters["scale_factor"] * np.sqrt(scale_range[i])
        target = max(0, 1.0 - thresholds[i])
        while float(ev.loss) > target:                                        # iterative_refinement_step :326
            n_refine += 1
            ev = WegEvaluation(denoiser, ev.latents, t, enc_t, masks_t, focus_indices, eot)
            if float(ev.loss) != 0:
                ev.latents = ev.latents - float(step_size) * ev.grad()        # update_latent
            if n_refine >= weg_parameters["max_refinement_steps"]:
                break
        ev = WegEvaluation(denoiser, ev.latents, t, enc_t, masks_t, focus_indices, eot)   # :368-387
    latents = ev.latents
    if i < weg_parameters["max_iter_to_alter"] and float(ev.loss) != 0:       # :490-495
        latents = latents - float(weg_parameters["scale_factor"] * np.sqrt(scale_range[i])) * ev.grad()
    if log is not None:
        log.append({"loss": float(ev.loss), "n_refine": n_refine})
    return latents, scale_range
