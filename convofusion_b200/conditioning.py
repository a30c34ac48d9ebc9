"""Conditioning projections and the 7-branch guidance batch.

Mirrors of the reference modules, parameter names included:
  AudioConvEncoder        convofusion/models/architectures/audioenc.py:9-34
  T5TextEncoder.projection convofusion/models/architectures/t5.py:48-49,57   (the frozen T5 body is out of
                          scope: callers pass its last hidden state [N, Lt, 768] + attention mask)
  TextAudioController     audioenc.py:37-91
  TextAudioMotionFuser    convofusion/models/architectures/condfuser.py:8-51

`guidance_memory()` builds the conditioning ONCE per clip in de-duplicated form: across the reference's
seven guidance branches (convofusion.py:909-929) every stream takes only two values -- the clip's own
conditioning or the batch-wide unconditional constant -- so memory holds 1 + B slots per stream and a
[7*B] slot table per stream says which one each (branch, clip) attends to.
"""
from __future__ import annotations

import functools
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

from . import _lib

# branch order: all_drop, text, audio, spk, apb, lsnid, full (convofusion.py:910); the stream each
# single-modality branch keeps conditional, as an index into (spkemb, alsn, tlsn, apb, lsnemb)
BRANCH_STREAM = {1: 2, 2: 1, 3: 0, 4: 3, 5: 4}


def _linear(x: Tensor, weight: Tensor, bias: Tensor, act: str = "none", a_act: str = "none") -> Tensor:
    if x.device.type != "cuda":
        raise _lib.CfbError("conditioning projections need CUDA tensors: convofusion_b200 has no CPU path")
    lead = x.shape[:-1]
    x2 = x.detach().to(torch.float32).reshape(-1, x.shape[-1]).contiguous()
    w = weight.detach().to(torch.float32).contiguous()
    b = bias.detach().to(torch.float32).contiguous()
    out = torch.empty(x2.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().cfb_linear(x2.data_ptr(), 0, w.data_ptr(), b.data_ptr(), out.data_ptr(), 0, x2.shape[0],
                                         w.shape[0], w.shape[1], _lib.ACT[act], _lib.ACT[a_act], 0, _lib.GEMM_SIMT,
                                         _lib.stream_ptr()))
    return out.reshape(*lead, w.shape[0])


class AudioConvEncoder(nn.Module):
    def __init__(self, input_size=80, hidden_size=256, latent_dim=512, **kwargs):
        super().__init__()
        # Sequential indices match the reference: 0 Linear, 1 Dropout, 2 LeakyReLU, 3 Linear, 4 Dropout, 5 LeakyReLU
        self.main = nn.Sequential(nn.Linear(input_size, hidden_size), nn.Identity(), nn.Identity(),
                                  nn.Linear(hidden_size, latent_dim), nn.Identity(), nn.Identity())
        self.out_net = nn.Linear(latent_dim, latent_dim)
        self.max_seq_len, self.fps = kwargs.get("max_seq_len", 128), kwargs.get("fps", 25)
        self.sample_rate, self.hop_length = kwargs.get("sample_rate", 16000), kwargs.get("hop_length", 512)
        self.audio_max_length = int((self.max_seq_len / self.fps) * self.sample_rate // self.hop_length + 1)

    def forward(self, inputs: Tensor) -> Tensor:
        if inputs.device.type != "cuda":
            raise _lib.CfbError("AudioConvEncoder needs CUDA tensors: convofusion_b200 has no CPU path")
        lead = inputs.shape[:-1]
        x = inputs.detach().to(torch.float32).reshape(-1, inputs.shape[-1]).contiguous()
        rows = x.shape[0]
        l0, l1, l2 = self.main[0], self.main[3], self.out_net
        p = [t.detach().to(torch.float32).contiguous() for t in (l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias)]
        hid, dout = l0.weight.shape[0], l2.weight.shape[0]
        t0 = torch.empty(rows, hid, device=x.device)
        t1 = torch.empty(rows, dout, device=x.device)
        out = torch.empty(rows, dout, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().cfb_audio_encoder(x.data_ptr(), rows, *[t.data_ptr() for t in p], x.shape[1], hid,
                                                    dout, t0.data_ptr(), t1.data_ptr(), out.data_ptr(),
                                                    _lib.stream_ptr()))
        return out.reshape(*lead, dout)


class T5TextEncoder(nn.Module):
    """Projection half of the reference's T5TextEncoder; takes the T5 last hidden state, not strings."""

    def __init__(self, latent_dim=512, encoded_dim=768, text_max_length=200, **kwargs):
        super().__init__()
        self.latent_dim, self.text_max_length = latent_dim, text_max_length
        self.projection = nn.Sequential(nn.Identity(), nn.Linear(encoded_dim, latent_dim))   # (ReLU, Linear)

    def forward(self, t5_hidden: Tensor, attention_mask: Tensor, return_map: bool = False):
        lin = self.projection[1]
        return _linear(t5_hidden, lin.weight, lin.bias, a_act="relu"), attention_mask, None


class TextAudioController(nn.Module):
    def __init__(self, out_dim=512, text_encoder: Optional[nn.Module] = None, audio_encoder: Optional[nn.Module] = None):
        super().__init__()
        self.text_encoder = text_encoder if text_encoder is not None else T5TextEncoder(out_dim)
        self.audio_encoder = audio_encoder if audio_encoder is not None else AudioConvEncoder(latent_dim=out_dim)
        self.out_dim = out_dim
        # unused on the 'spk' / 'lsn' paths the sampler takes, kept for state_dict fidelity
        self.text_time_proj = nn.Linear(self.text_encoder.text_max_length, out_dim)
        self.audio_time_proj = nn.Linear(self.audio_encoder.audio_max_length, out_dim)
        self.out_net = nn.Linear(out_dim, out_dim)

    def forward(self, text_hidden: Tensor, text_attention_mask: Tensor, audio: Tensor, person_type: str = "lsn",
                return_textmap: bool = False):
        if person_type == "spk-ta":
            raise NotImplementedError("person_type='spk-ta' is not on the sampling path (convofusion.py:932-934)")
        text_emb, mask, tmap = self.text_encoder(text_hidden, text_attention_mask, return_map=return_textmap)
        text_mask = ~mask.bool()                 # audioenc.py:61: True = padding
        audio_emb = self.audio_encoder(audio)
        return audio_emb, text_emb, None, text_mask, tmap, None


class TextAudioMotionFuser(nn.Module):
    def __init__(self, out_dim=512, latent_dim=128):
        super().__init__()
        self.out_dim = out_dim
        self.active_passive_emb = nn.Embedding(3, out_dim)
        self.lsn_id_emb = nn.Embedding(5 + 1 + 30, out_dim)
        self.latent_proj = nn.Sequential(nn.Linear(latent_dim, 128), nn.Identity(), nn.Linear(128, out_dim), nn.Identity())

    def forward(self, spkemb, alsn, tlsn, active_passive_bit, lsn_id):
        # condfuser.py:41-51: two table lookups (pure gathers, no arithmetic)
        apb = self.active_passive_emb.weight[active_passive_bit.to(torch.long)]
        ids = torch.as_tensor(lsn_id, dtype=torch.long, device=spkemb.device)
        lsnemb = self.lsn_id_emb.weight[ids].unsqueeze(1)
        return spkemb, alsn, tlsn, apb, lsnemb


def uncond_mel(n_frames: int, n_mel: int, device) -> Tensor:
    """convofusion.py:914-915."""
    m = -90 * torch.ones(1, n_frames, n_mel, device=device)
    m[..., 40:45] = 0
    return m


def guidance_branches(return_attention: bool = False, spk_is_uncond: bool = False) -> List[int]:
    """Branches of convofusion.py:910 that have to be evaluated.  Branch 6 (full-cond) has guidance weight 0
    (convofusion.py:539) and only matters for its attention maps; branch 3 (speaker-only) equals branch 0 when the
    speaker stream of every clip IS the unconditional prompt (monadic BEAT clips, dataset.py:185-199), so its term
    guidance_scale * (e_3 - e_0) is an exact zero."""
    br = [0, 1, 2] + ([] if spk_is_uncond else [3]) + [4, 5]
    return br + [6] if return_attention else br


@functools.lru_cache(maxsize=64)
def _slots_host(n_clips: int, branches: Tuple[int, ...]) -> Tuple[Tensor, ...]:
    b = torch.arange(n_clips, dtype=torch.int32)
    out = []
    for x in range(5):
        rows = [b + 1 if (g == 6 or BRANCH_STREAM.get(g) == x) else torch.zeros_like(b) for g in branches]
        out.append(torch.cat(rows))
    return tuple(out)


@functools.lru_cache(maxsize=64)
def _slots_device(n_clips: int, branches: Tuple[int, ...], device: str) -> Tuple[Tensor, ...]:
    return tuple(t.to(device) for t in _slots_host(n_clips, branches))


def guidance_slots(n_clips: int, branches, device) -> List[Tensor]:
    """slot[x][i*B + b] for the i-th evaluated branch g = branches[i]: 0 = unconditional constant, 1 + b = clip b's
    own conditioning.  `branches` is a list of branch ids or an int n (= branches 0..n-1).  The
    tables are a pure function of (B, branches): built once and cached per device (read-only: do not modify)."""
    br = tuple(range(branches)) if isinstance(branches, int) else tuple(int(g) for g in branches)
    dev = torch.device(device)
    if dev.type == "cpu":
        return list(_slots_host(n_clips, br))
    return list(_slots_device(n_clips, br, str(dev)))


def guidance_slots_host(n_clips: int, branches) -> List[Tensor]:
    """Host copy of `guidance_slots` (int32, CPU): lets the library derive its execution plan without reading the
    device tables back (cfb_memory.slot_host)."""
    br = tuple(range(branches)) if isinstance(branches, int) else tuple(int(g) for g in branches)
    return list(_slots_host(n_clips, br))


def guidance_memory(controller: TextAudioController, fuser: TextAudioMotionFuser, clip: Dict[str, Tensor],
                    uncond_text: Tensor, uncond_text_attn: Tensor):
    """De-duplicated conditioning for a batch of clips.

    clip: text_lsn [B,Lt,768] / text_lsn_attn [B,Lt] (1 = token), text_spk / text_spk_attn likewise,
    mel_lsn [B,161,80], apb [B,8] int, lsn_id list[int].  uncond_text [Lt,768], uncond_text_attn [Lt] are the
    T5 features of the reference's '-'*10 prompt.  Returns (enc: 5 x [1+B, M_x, 512], masks dict)."""
    dev = clip["mel_lsn"].device
    B = clip["mel_lsn"].shape[0]
    u, ua = uncond_text.unsqueeze(0), uncond_text_attn.unsqueeze(0)
    t_lsn = torch.cat([u, clip["text_lsn"]]); a_lsn = torch.cat([ua, clip["text_lsn_attn"]])
    t_spk = torch.cat([u, clip["text_spk"]]); a_spk = torch.cat([ua, clip["text_spk_attn"]])
    mel = torch.cat([uncond_mel(clip["mel_lsn"].shape[1], clip["mel_lsn"].shape[2], dev), clip["mel_lsn"]])
    alsn, tlsn, _, tl_mask, _, _ = controller(t_lsn, a_lsn, mel, person_type="lsn")
    tspk, ts_attn, _ = controller.text_encoder(t_spk, a_spk)     # speaker audio is encoded then dropped (:933,970)
    ts_mask = ~ts_attn.bool()
    apb = torch.cat([2 * torch.ones_like(clip["apb"][:1]), clip["apb"]])
    ids = [0] + list(clip["lsn_id"])
    enc = fuser(tspk, alsn, tlsn, apb, ids)
    return enc, {"alsn": None, "tlsn": tl_mask, "spkemb": ts_mask}


def expand_guidance_batch(enc: Sequence[Tensor], masks: Dict[str, Optional[Tensor]], n_clips: int, n_branch: int = 7):
    """The reference's as-written [7*B, ...] batch (convofusion.py:909-929), gathered from the slots."""
    slots = guidance_slots(n_clips, n_branch, enc[0].device)
    enc7 = [e[s.long()] for e, s in zip(enc, slots)]
    names = ("spkemb", "alsn", "tlsn", "apb", "lsnemb")
    masks7 = {n: (masks[n][s.long()] if masks.get(n) is not None else None) for n, s in zip(names, slots)
              if n in masks}
    return enc7, masks7
