"""Seeded synthetic weights and inputs: there is no network for checkpoints or data, so
tests, goldens and the benchmark all draw from these generators.  Pure torch-CPU RNG -> reproducible on any box
with the same torch build."""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn


@torch.no_grad()
def randomize_(module: nn.Module, seed: int = 1234) -> nn.Module:
    """Deterministic non-degenerate init: every layer distinct, biases and LayerNorm affine non-trivial (the
    reference's deepcopy-cloned layers start identical, which would hide layer-indexing bugs)."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        leaf = name.split(".")[-1]
        is_norm = "norm" in name.split(".")[-2] if "." in name else False
        if p.dim() >= 2 and "embedding" not in name and "emb." not in name and "token" not in name:
            fan_in = p.shape[-1]
            v = torch.randn(p.shape, generator=g) / fan_in ** 0.5
        elif is_norm and leaf == "weight":
            v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
        elif is_norm and leaf == "bias":
            v = 0.1 * torch.randn(p.shape, generator=g)
        elif leaf == "bias" or leaf == "in_proj_bias":
            v = 0.05 * torch.randn(p.shape, generator=g)
        elif name == "cond_params":
            v = torch.full(p.shape, 0.2)
        else:   # embedding tables, global motion tokens
            v = 0.5 * torch.randn(p.shape, generator=g)
        p.copy_(v.to(p.dtype))
    return module


def synthetic_clip(n_clips: int, seed: int = 1234, dyadic: bool = False, text_len: int = 32, n_mel_frames: int = 161
                   ) -> Dict[str, object]:
    """One batch of featurised clips at BEAT shapes.  Text is the T5 last hidden state."""
    g = torch.Generator().manual_seed(seed)
    B, Lt = n_clips, text_len
    clip = {
        "mel_lsn": torch.rand(B, n_mel_frames, 80, generator=g) * 80.0 - 80.0,
        "text_lsn": torch.randn(B, Lt, 768, generator=g),
        "apb": torch.randint(0, 2, (B, 8), generator=g),
        "lsn_id": [int(v) for v in torch.randint(1, 36, (B,), generator=g)],
    }
    lsn_valid = torch.randint(12, Lt + 1, (B,), generator=g) if B > 1 else torch.tensor([20])
    clip["text_lsn_attn"] = (torch.arange(Lt)[None, :] < lsn_valid[:, None]).long()
    uncond_text = torch.randn(Lt, 768, generator=g)
    uncond_attn = (torch.arange(Lt) < 4).long()
    if dyadic:   # DnD: the speaker's transcript is real
        clip["text_spk"] = torch.randn(B, Lt, 768, generator=g)
        spk_valid = torch.randint(8, Lt + 1, (B,), generator=g)
        clip["text_spk_attn"] = (torch.arange(Lt)[None, :] < spk_valid[:, None]).long()
        clip["spk_is_uncond"] = False
    else:        # monadic BEAT: speaker stream is the unconditional prompt (dataset.py:185-199)
        clip["text_spk"] = uncond_text.unsqueeze(0).repeat(B, 1, 1)
        clip["text_spk_attn"] = uncond_attn.unsqueeze(0).repeat(B, 1)
        clip["spk_is_uncond"] = True      # the data loader knows (dataset.py:185-199); see ConvoFusionSampler.sample
    return {"clip": clip, "uncond_text": uncond_text, "uncond_text_attn": uncond_attn}


def to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: to_device(v, device) for k, v in obj.items()}
    return obj


def synthetic_text_features(text: str, text_len: int = 32):
    """Stand-in for the frozen T5 body on an arbitrary string: (last hidden state [Lt,768], attention mask [Lt]),
    a pure function of the string (seeded by its CRC32); the valid length grows with the word count."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(text.encode("utf-8")))
    hid = torch.randn(text_len, 768, generator=g)
    valid = min(text_len, 2 + len(text.split()))
    return hid, (torch.arange(text_len) < valid).long()


def synthetic_long_batch(n_streams: int, n_parts: int, seed: int = 1234) -> Dict[str, object]:
    """A dataloader batch in the layout `process_samples` consumes (unbounded_synthesis.py:244-271): n_parts * 128
    frames per stream, mel / active-passive bits / audio for the whole span, and word-level transcripts with
    timestamps (`seg_*`: [((start, end), word), ...] or the unconditional prompt).  Stream 0's speaker is silent."""
    g = torch.Generator().manual_seed(seed)
    B, T = n_streams, n_parts * 128
    words = lambda tag, b: [((0.5 * i, 0.5 * i + 0.45), f"{tag}{b}w{i}") for i in range(int(T / 25 / 0.5))]
    return {
        "length": [T] * B,
        "motion_lsn": torch.randn(B, T, 189, generator=g),
        "motion_spk": torch.randn(B, T, 189, generator=g),
        "melspec_lsn": torch.rand(B, n_parts * 160 + 1, 80, generator=g) * 80.0 - 80.0,
        "melspec_spk": torch.rand(B, n_parts * 160 + 1, 80, generator=g) * 80.0 - 80.0,
        "active_passive_lsn": torch.randint(0, 2, (B, n_parts * 8), generator=g),
        "lsn_id": [int(v) for v in torch.randint(1, 36, (B,), generator=g)],
        "audio_lsn": torch.zeros(B, n_parts * 1000), "audio_spk": torch.zeros(B, n_parts * 1000),
        "combined_audio": torch.zeros(B, n_parts * 1000),
        "seg_lsn": [words("l", b) for b in range(B)],
        "seg_spk": ["-" * 10 if b == 0 else words("s", b) for b in range(B)],
        "text_lsn": [" ".join(w for _, w in words("l", b)) for b in range(B)],
        "text_spk": [" ".join(w for _, w in words("s", b)) for b in range(B)],
        "name": [f"stream{b}" for b in range(B)], "spk_name": ["spk"] * B, "lsn_name": ["lsn"] * B,
    }
