"""CPU oracle for the ConvoFusion sampling hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU, state_dict-driven restatement of the reference's
sampling path.  It is imported only by tests/, __graft_entry__.smoke() and the
baseline legs of bench.py (cpu_baseline / --impl reference on the host cores;
gpu_eager_baseline = the same eager torch code with its tensors on the B200, the
"PyTorch eager on the same box" yardstick of SURVEY 8d) -- never by the product package
(convofusion_b200/), which must fail loudly when its CUDA library is missing.

Pinning status
--------------
* Denoiser.forward, ConvoFusionVae.decode, AudioConvEncoder, the T5 projection and
  the condition fuser are PINNED: tools/make_golden.py runs the unmodified reference
  modules from /root/reference on seeded weights/inputs and tests/test_oracle.py
  checks this restatement against those committed outputs (tests/golden/*.pt).
* The 7-branch batch assembly, the two reverse loops (_diffusion_reverse /
  diffusion_reverse_forecast), the generation call chain (test_diffusion_forward) and
  the window driver (process_samples) are restated line by line AND PINNED to the
  reference's own code: tools/pin_reference_loops.py imports those functions unmodified
  (empty stand-in modules for the absent lightning / torchmetrics / kornia / nltk ...
  packages they never call on this path, a stand-in for the frozen T5 body), runs them
  with the reference Denoiser / VAE / encoders and asserts that the functions below
  reproduce them bit for bit; tests/golden/ref_loops.pt + tests/test_oracle.py keep it so.
* diffusers==0.14.0 (DDIM/DDPM schedulers; reference environment.yml:85) is absent
  from /root/reference and from this image: its published algorithm is restated
  below.  PARITY UNPINNED for the scheduler arithmetic.

Every function cites the reference file:line it follows (paths relative to the
reference repository root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
STREAMS = ("spkemb", "alsn", "tlsn", "apb", "lsnemb")  # cross_attention.py:579 memory order
N_BRANCH = 7  # convofusion.py:60 clf_guidance_drops + 1


# --------------------------------------------------------------------------- basics
def sine_pe(max_len: int, d_model: int, dtype=torch.float32) -> Tensor:
    """position_encoding.py:119-125 -> [max_len, d_model] (the buffer is [max_len,1,d])."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


def timestep_embedding(timesteps: Tensor, dim: int, flip_sin_to_cos: bool = True,
                       freq_shift: float = 0.0) -> Tensor:
    """embeddings.py:245-285 (scale=1, max_period=10000)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def lengths_to_mask(lengths: Sequence[int], max_len: Optional[int] = None) -> Tensor:
    """utils/temos_utils.py:11-18."""
    lengths = torch.tensor(list(lengths))
    max_len = max_len if max_len else int(lengths.max())
    return torch.arange(max_len).expand(len(lengths), max_len) < lengths.unsqueeze(1)


def mha(query: Tensor, key: Tensor, value: Tensor, in_w: Tensor, in_b: Tensor,
        out_w: Tensor, out_b: Tensor, nheads: int,
        key_padding_mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """torch.nn.MultiheadAttention as called at cross_attention.py:370-377,570,593-626:
    seq-first [L,B,E], packed in_proj (rows 0:E q, E:2E k, 2E:3E v), q scaled by
    1/sqrt(head_dim), additive -inf key padding mask, softmax, out_proj; returns the
    head-averaged weights [B,L,S] (need_weights=True default)."""
    L, B, E = query.shape
    S = key.shape[0]
    hd = E // nheads
    q = F.linear(query, in_w[:E], in_b[:E])
    k = F.linear(key, in_w[E:2 * E], in_b[E:2 * E])
    v = F.linear(value, in_w[2 * E:], in_b[2 * E:])
    q = q.reshape(L, B * nheads, hd).transpose(0, 1)
    k = k.reshape(S, B * nheads, hd).transpose(0, 1)
    v = v.reshape(S, B * nheads, hd).transpose(0, 1)
    scores = torch.bmm(q * math.sqrt(1.0 / hd), k.transpose(1, 2))  # [B*H, L, S]
    if key_padding_mask is not None:
        m = torch.zeros(B, 1, 1, S, dtype=scores.dtype, device=scores.device)
        m = m.masked_fill(key_padding_mask.view(B, 1, 1, S), float("-inf"))
        scores = (scores.view(B, nheads, L, S) + m).view(B * nheads, L, S)
    attn = torch.softmax(scores, dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).reshape(L, B, E)
    out = F.linear(out, out_w, out_b)
    return out, attn.view(B, nheads, L, S).mean(dim=1)


def _mha_sd(sd, pfx, q, k, v, nheads, kpm=None):
    return mha(q, k, v, sd[pfx + "in_proj_weight"], sd[pfx + "in_proj_bias"],
               sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"], nheads, kpm)


def _lin(sd, pfx, x):
    return F.linear(x, sd[pfx + "weight"], sd[pfx + "bias"])


def _ln(sd, pfx, x):
    return layer_norm(x, sd[pfx + "weight"], sd[pfx + "bias"])


# --------------------------------------------------------------------------- denoiser
def time_block(sd, pfx: str, h: Tensor, emb: Tensor) -> Tensor:
    """cross_attention.py:426-439 (dropout off)."""
    emb_out = _lin(sd, pfx + "emb_layers.1.", F.silu(emb))
    scale, shift = torch.chunk(emb_out, 2, dim=2)
    h = _ln(sd, pfx + "norm.", h) * (1 + scale) + shift
    return _lin(sd, pfx + "out_layers.2.", F.silu(h))


def decoder_layer_2att(sd, pfx: str, tgt: Tensor, memory: Sequence[Tensor], time_embed: Tensor,
                       masks: Dict[str, Optional[Tensor]], nheads: int):
    """cross_attention.py:556-664 TransformerDecoderLayer2Att.forward_pre (pos/query_pos None)."""
    t2 = _ln(sd, pfx + "norm1.", tgt)
    t2, _ = _mha_sd(sd, pfx + "self_attn.", t2, t2, t2, nheads)
    tgt = tgt + t2
    tgt = tgt + time_block(sd, pfx + "time_block1.", tgt, time_embed)
    t2 = _ln(sd, pfx + "norm2.", tgt)
    outs, atts = [], []
    for name, mem in zip(STREAMS, memory):
        m = _ln(sd, pfx + f"{name}_norm.", mem)
        o, a = _mha_sd(sd, pfx + f"multihead_attn_{name}.", t2, m, m, 1, masks.get(name))
        outs.append(o)
        atts.append(a)
    tgt = tgt + _lin(sd, pfx + "att_fuser.", torch.cat(outs, dim=-1))
    tgt = tgt + time_block(sd, pfx + "time_block2.", tgt, time_embed)
    t2 = _ln(sd, pfx + "norm3.", tgt)
    t2 = _lin(sd, pfx + "linear2.", F.gelu(_lin(sd, pfx + "linear1.", t2)))
    return tgt + t2, atts


def denoiser_forward(sd: Dict[str, Tensor], sample: Tensor, timestep, enc: Sequence[Tensor],
                     mem_mask_dict: Dict[str, Optional[Tensor]], prefix: str = "",
                     num_layers: int = 9, num_heads: int = 4,
                     return_hidden: bool = False):
    """denoiser.py:173-386 with arch=trans_dec, condition=text+audio, abl_plus=True.

    sample [BG,16,latent]; enc = (spk_emb, alsn, tlsn, apb, lsnemb) batch-first [BG,M_x,512];
    masks bool [BG,M_x] (True = ignore) under keys 'spkemb','alsn','tlsn'.
    Returns (eps [BG,16,latent], 5 x att [BG,num_layers,16,M_x])."""
    p = prefix
    d = sd[p + "latent_embd.weight"].shape[0]
    x = sample.permute(1, 0, 2)
    x = _lin(sd, p + "latent_embd.", x)                                   # :187
    bs = x.shape[1]
    t = torch.as_tensor(timestep).to(x.device).reshape(-1)[:1].expand(bs)
    temb = timestep_embedding(t, d, True, 0.0).to(x.dtype)                # :195-197
    temb = _lin(sd, p + "time_embedding.linear_2.",
                F.silu(_lin(sd, p + "time_embedding.linear_1.", temb))).unsqueeze(0)  # :199
    mems = [e.permute(1, 0, 2) + temb for e in enc]                       # :223-261
    x = x.clone()
    bh = sd[p + "bh_embedding.weight"]
    x[0::2] = x[0::2] + bh[0]                                             # :316-324
    x[1::2] = x[1::2] + bh[1]
    pe_q = sd[p + "query_pos.pe"][:, 0]
    x[0::2] = x[0::2] + pe_q[: x.shape[0] // 2, None]                    # position_encoding.py:160-161
    x[1::2] = x[1::2] + pe_q[: x.shape[0] // 2, None]
    pe_m = sd[p + "mem_pos.pe"][:, 0]
    ce = sd[p + "condition_embedding.weight"]
    mems = [m + ce[i] + pe_m[: m.shape[0], None] for i, m in enumerate(mems)]  # :332-353
    masks = {"spkemb": mem_mask_dict.get("spkemb"), "alsn": mem_mask_dict.get("alsn"),
             "tlsn": mem_mask_dict.get("tlsn"), "apb": mem_mask_dict.get("apb"),
             "lsnemb": mem_mask_dict.get("lsnemb")}
    per_layer = []
    for l in range(num_layers):                                           # cross_attention.py:217-228
        x, atts = decoder_layer_2att(sd, p + f"decoder.layers.{l}.", x, mems, temb, masks, num_heads)
        per_layer.append(atts)
    att_mats = [torch.stack([pl[i] for pl in per_layer]).permute(1, 0, 2, 3) for i in range(5)]  # :234
    hidden = x
    x = _ln(sd, p + "decoder.norm.", x)                                   # :238-239
    x = _lin(sd, p + "latent_proj.", x)                                   # denoiser.py:382
    out = x.permute(1, 0, 2)
    if return_hidden:
        return out, att_mats, hidden
    return out, att_mats


# --------------------------------------------------------------------------- VAE decode
def vae_decoder_layer(sd, pfx, tgt, memory, tgt_kpm, nheads):
    """cross_attention.py:361-382 TransformerDecoderLayer.forward_pre (pos None)."""
    t2 = _ln(sd, pfx + "norm1.", tgt)
    t2, _ = _mha_sd(sd, pfx + "self_attn.", t2, t2, t2, nheads, tgt_kpm)
    tgt = tgt + t2
    t2 = _ln(sd, pfx + "norm2.", tgt)
    t2, _ = _mha_sd(sd, pfx + "multihead_attn.", t2, memory, memory, nheads)
    tgt = tgt + t2
    t2 = _ln(sd, pfx + "norm3.", tgt)
    t2 = _lin(sd, pfx + "linear2.", F.gelu(_lin(sd, pfx + "linear1.", t2)))
    return tgt + t2


def skip_decoder(sd, pfx, tgt, memory, tgt_kpm, nheads, num_layers=5):
    """cross_attention.py:89-125 SkipTransformerDecoder.forward."""
    nb = (num_layers - 1) // 2
    x, xs = tgt, []
    for i in range(nb):
        x = vae_decoder_layer(sd, pfx + f"input_blocks.{i}.", x, memory, tgt_kpm, nheads)
        xs.append(x)
    x = vae_decoder_layer(sd, pfx + "middle_block.", x, memory, tgt_kpm, nheads)
    for i in range(nb):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = _lin(sd, pfx + f"linear_blocks.{i}.", x)
        x = vae_decoder_layer(sd, pfx + f"output_blocks.{i}.", x, memory, tgt_kpm, nheads)
    return _ln(sd, pfx + "norm.", x)


def vae_decode(sd: Dict[str, Tensor], z: Tensor, lengths: Sequence[int], prefix: str = "",
               num_layers: int = 5, num_heads: int = 2) -> Tensor:
    """vae.py:268-372 (arch=encoder_decoder, pe_type=convofusion).
    z [2, B, n_chunks, latent]; returns [B, nframes, 189]."""
    p = prefix
    mask = lengths_to_mask(lengths).to(z.device)
    bs, nframes = mask.shape
    dlat = z.shape[-1]
    pe_q = sd[p + "query_pos_decoder.pe"][:, 0]
    pe_m = sd[p + "mem_pos_decoder.pe"][:, 0]
    queries = torch.zeros(nframes, bs, dlat, dtype=z.dtype, device=z.device) + pe_q[:nframes, None]   # :277,321
    outs = []
    for part, zz in zip(("body", "hands"), torch.chunk(z, 2, dim=0)):
        m = zz.squeeze(0).permute(1, 0, 2)                                           # :279-285
        m = m + pe_m[: m.shape[0], None]                                             # :322,331
        o = skip_decoder(sd, p + f"{part}_decoder.", queries, m, ~mask, num_heads, num_layers)
        outs.append(_lin(sd, p + f"{part}_final_layer.", o))                          # :352-353
    output = torch.cat(outs, dim=-1)
    output[~mask.T] = 0                                                               # :362
    return output.permute(1, 0, 2)


def vae_encoder_layer(sd, pfx, src, kpm, nheads):
    """cross_attention.py:288-300 TransformerEncoderLayer.forward_pre (pos None)."""
    s2 = _ln(sd, pfx + "norm1.", src)
    s2, _ = _mha_sd(sd, pfx + "self_attn.", s2, s2, s2, nheads, kpm)
    src = src + s2
    s2 = _ln(sd, pfx + "norm2.", src)
    s2 = _lin(sd, pfx + "linear2.", F.gelu(_lin(sd, pfx + "linear1.", s2)))
    return src + s2


def skip_encoder(sd, pfx, src, kpm, nheads, num_layers=5):
    """cross_attention.py:41-64 SkipTransformerEncoder.forward."""
    nb = (num_layers - 1) // 2
    x, xs = src, []
    for i in range(nb):
        x = vae_encoder_layer(sd, pfx + f"input_blocks.{i}.", x, kpm, nheads)
        xs.append(x)
    x = vae_encoder_layer(sd, pfx + "middle_block.", x, kpm, nheads)
    for i in range(nb):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = _lin(sd, pfx + f"linear_blocks.{i}.", x)
        x = vae_encoder_layer(sd, pfx + f"output_blocks.{i}.", x, kpm, nheads)
    return _ln(sd, pfx + "norm.", x)


def vae_encode(sd: Dict[str, Tensor], features: Tensor, lengths: Sequence[int], prefix: str = "",
               num_layers: int = 5, num_heads: int = 2):
    """vae.py:162-258 (MLP_DIST False, pe_type convofusion) up to the distribution parameters:
    returns (mu, std [2, B*T/16, d], chunk-root-subtracted features [B, T, 189])."""
    p = prefix
    bs, nframes, _ = features.shape
    mask = lengths_to_mask(lengths, nframes)
    n_chunks = nframes // 16
    mf = features.clone().reshape(bs * n_chunks, 16, -1)
    root = mf[:, :1, :3] * torch.tensor([1.0, 0.0, 1.0], dtype=features.dtype)          # :182-184
    mf[:, :, :3] = mf[:, :, :3] - root
    mask = mask.reshape(bs * n_chunks, 16)
    n = bs * n_chunks
    pe = sd[p + "query_pos_encoder.pe"][:, 0]
    mus, lvs = [], []
    nb_feats = sd[p + "body_skel_embedding.weight"].shape[1]
    for part, xs_ in (("body", mf[:, :, :nb_feats]), ("hands", mf[:, :, nb_feats:])):
        x = _lin(sd, p + f"{part}_skel_embedding.", xs_).permute(1, 0, 2)               # [16, n, d]
        tok = sd[p + f"{part}_global_motion_token"]
        dist = torch.tile(tok[:, None, :], (1, n, 1))                                   # :204-205
        aug = torch.cat((torch.ones(n, tok.shape[0], dtype=torch.bool), mask), 1)        # :208-216
        xseq = torch.cat((dist, x), 0)
        xseq = xseq + pe[: xseq.shape[0], None]                                         # :224
        out = skip_encoder(sd, p + f"{part}_encoder.", xseq, ~aug, num_heads, num_layers)[: tok.shape[0]]
        mus.append(out[0:1])
        lvs.append(out[1:2])
    mu, logvar = torch.cat(mus, 0), torch.cat(lvs, 0)                                   # :250-257
    std = logvar.exp().pow(0.5)                                                          # :260
    return mu, std, mf.reshape(bs, nframes, -1)


def feats_to_keypoints3d(feats: Tensor) -> Tensor:
    """models/modeltype/base.py:204-209 (numpy there): [T, 189] -> [T, 63, 3], in-place order preserved."""
    p = feats.reshape(-1, 63, 3).clone()
    p = p / 3
    p[:, 43:, :] = p[:, 43:, :] + p[:, [11], :]
    p[:, 23:43, :] = p[:, 23:43, :] + p[:, [7], :]
    p[:, 1:, :] = p[:, 1:, :] + p[:, :1, :]
    return p


# --------------------------------------------------------------------------- conditioning
def audio_encoder(sd, mel: Tensor, prefix: str = "text_audio_encoder.audio_encoder.") -> Tensor:
    """audioenc.py:13-34: Linear -> LeakyReLU(0.1) -> Linear -> LeakyReLU(0.1) -> out_net."""
    h = F.leaky_relu(_lin(sd, prefix + "main.0.", mel), 0.1)
    h = F.leaky_relu(_lin(sd, prefix + "main.3.", h), 0.1)
    return _lin(sd, prefix + "out_net.", h)


def text_projection(sd, t5_hidden: Tensor,
                    prefix: str = "text_audio_encoder.text_encoder.projection.1.") -> Tensor:
    """t5.py:48-49,57: Sequential(ReLU, Linear(768,512)) on the T5 last hidden state."""
    return _lin(sd, prefix, F.relu(t5_hidden))


def condition_fuser(sd, apb: Tensor, lsn_id: Sequence[int], prefix: str = "condition_fuser."):
    """condfuser.py:32-51: two embedding lookups."""
    a = sd[prefix + "active_passive_emb.weight"][apb.long()]
    l = sd[prefix + "lsn_id_emb.weight"][torch.as_tensor(list(lsn_id), device=apb.device).long()].unsqueeze(1)
    return a, l


def assemble_guidance_batch(sd, clip: Dict[str, Tensor], uncond_text: Tensor, uncond_text_mask: Tensor):
    """convofusion.py:909-973 on synthetic features.

    clip: text_lsn [B,Lt,768], text_lsn_mask [B,Lt] (True = pad), text_spk/text_spk_mask,
    mel_lsn [B,161,80], apb [B,8] int, lsn_id list[int].  uncond_text [Lt,768] is the T5
    hidden state of the reference's '-'*10 prompt (synthetic here), mask [Lt].
    Branch order: all_drop, text, audio, spk, apb, lsnid, full."""
    B = clip["mel_lsn"].shape[0]
    U = uncond_text.unsqueeze(0).expand(B, -1, -1)
    Um = uncond_text_mask.unsqueeze(0).expand(B, -1)
    tl, tlm = clip["text_lsn"], clip["text_lsn_mask"]
    ts, tsm = clip["text_spk"], clip["text_spk_mask"]
    text_lsn = torch.cat([U, tl, U, U, U, U, tl])
    text_lsn_mask = torch.cat([Um, tlm, Um, Um, Um, Um, tlm])
    text_spk = torch.cat([U, U, U, ts, U, U, ts])
    text_spk_mask = torch.cat([Um, Um, Um, tsm, Um, Um, tsm])
    mel = clip["mel_lsn"]
    uncond_mel = -90 * torch.ones_like(mel)
    uncond_mel[..., 40:45] = 0
    mel7 = torch.cat([uncond_mel, uncond_mel, mel, uncond_mel, uncond_mel, uncond_mel, mel])
    apb = clip["apb"]
    two = 2 * torch.ones_like(apb)
    apb7 = torch.cat([two, two, two, two, apb, two, apb])
    ids = list(clip["lsn_id"])
    ids7 = [0] * (5 * B) + ids + ids
    tspk = text_projection(sd, text_spk)
    tlsn = text_projection(sd, text_lsn)
    alsn = audio_encoder(sd, mel7)
    apb_e, id_e = condition_fuser(sd, apb7, ids7)
    enc = (tspk, alsn, tlsn, apb_e, id_e)
    masks = {"alsn": None, "tlsn": text_lsn_mask, "spkemb": text_spk_mask}
    return enc, masks


# --------------------------------------------------------------------------- schedulers
class _SchedOut:
    def __init__(self, prev_sample, pred_original_sample):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class _SchedBase:
    """Shared table build of diffusers 0.14.0 DDPMScheduler/DDIMScheduler.__init__
    (beta_schedule 'scaled_linear' / 'linear'), all in float32 torch like the original."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", clip_sample=True, prediction_type="epsilon", **kw):
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.clip_sample = clip_sample
        self.prediction_type = prediction_type
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def add_noise(self, original, noise, timesteps):
        t = torch.as_tensor(timesteps).reshape(-1).long()
        sa = self.alphas_cumprod[t] ** 0.5
        sb = (1 - self.alphas_cumprod[t]) ** 0.5
        while sa.dim() < original.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original + sb * noise

    def _x0(self, model_output, sample, a_t):
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        elif self.prediction_type == "sample":
            x0 = model_output
        else:
            raise NotImplementedError(self.prediction_type)
        if self.clip_sample:
            x0 = torch.clamp(x0, -1, 1)
        return x0


class DDIMSchedulerOracle(_SchedBase):
    """diffusers 0.14.0 DDIMScheduler (set_timesteps / step), restated; see SURVEY a13."""

    def __init__(self, set_alpha_to_one=True, steps_offset=0, **kw):
        super().__init__(**kw)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.steps_offset = steps_offset

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts) + self.steps_offset

    def step(self, model_output, timestep, sample, eta=0.0, variance_noise=None, generator=None):
        t = int(timestep)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = self._x0(model_output, sample, a_t)
        variance = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        std = eta * variance ** 0.5
        if self.prediction_type != "epsilon":
            # The shipped config predicts epsilon (config_cf_beatdnd.yaml:46); how 0.14.0's DDIM
            # forms the direction term for 'sample' cannot be checked offline, so refuse.
            raise NotImplementedError("DDIM oracle restates prediction_type='epsilon' only")
        direction = (1 - a_p - std ** 2) ** 0.5 * model_output
        prev_sample = a_p ** 0.5 * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev_sample = prev_sample + std * variance_noise
        return _SchedOut(prev_sample, x0)


class DDPMSchedulerOracle(_SchedBase):
    """diffusers 0.14.0 DDPMScheduler (variance_type fixed_small), restated; see SURVEY a13."""

    def __init__(self, variance_type="fixed_small", **kw):
        super().__init__(**kw)
        self.variance_type = variance_type

    def set_timesteps(self, n):
        n = min(self.num_train_timesteps, n)
        self.num_inference_steps = n
        ts = np.arange(0, self.num_train_timesteps, self.num_train_timesteps // n)[::-1].copy()
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output, timestep, sample, variance_noise=None, generator=None):
        t = int(timestep)
        n = self.num_inference_steps if self.num_inference_steps else self.num_train_timesteps
        prev = t - self.num_train_timesteps // n
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else torch.tensor(1.0)
        b_t, b_p = 1 - a_t, 1 - a_p
        cur_a = a_t / a_p
        cur_b = 1 - cur_a
        x0 = self._x0(model_output, sample, a_t)
        c0 = (a_p ** 0.5 * cur_b) / b_t
        c1 = cur_a ** 0.5 * b_p / b_t
        prev_sample = c0 * x0 + c1 * sample
        if t > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            var = torch.clamp(b_p / b_t * cur_b, min=1e-20)
            prev_sample = prev_sample + (var ** 0.5) * variance_noise
        return _SchedOut(prev_sample, x0)


# --------------------------------------------------------------------------- guidance + loops
def guidance_combine(noise_pred: Tensor, guidance_scale: float) -> Tensor:
    """convofusion.py:527-541, evaluated in the reference's association order."""
    e0, e_text, e_audio, e_spk, e_apb, e_id, e_full = noise_pred.chunk(N_BRANCH)
    s = guidance_scale
    n_text = s * 1 * (e_text - e0)
    n_audio = s * 1 * (e_audio - e0)
    n_spk = s * 1 * (e_spk - e0)
    n_apb = s * 1 * (e_apb - e0)
    n_id = s * 1 * (e_id - e0)
    n_all = s * 0 * (e_full - e0)
    return e0 + (n_text + n_audio + n_spk + n_apb + n_id + n_all)


def diffusion_reverse(denoise_fn, scheduler, enc, masks, init_latents: Tensor, num_steps: int,
                      guidance_scale: float = 7.5, eta: float = 0.0,
                      step_noise: Optional[Tensor] = None, record: Optional[list] = None,
                      focus_indices: Sequence[Sequence[int]] = (), weg: Optional[dict] = None, weg_log: Optional[list] = None):
    """convofusion.py:391-549; focus_indices=[] is WEG off, otherwise every guided step is preceded by weg_pre_step
    (needs autograd: call outside torch.no_grad()).

    init_latents replaces torch.randn at :412 so the noise is shared with the CUDA path;
    step_noise [num_steps,B,16,latent] replaces the scheduler's internal randn (DDPM / eta>0).
    denoise_fn(sample, t, enc, masks) -> (eps, att_mats).  Returns (latents [16,B,latent],
    {t: 5 att maps of the full-cond branch})."""
    latents = init_latents * scheduler.init_noise_sigma                     # :419
    scheduler.set_timesteps(num_steps)                                      # :421
    att = {}
    scale_range = weg["scale_range"] if weg else None                        # :395
    for i, t in enumerate(scheduler.timesteps):                             # :435
        if len(focus_indices) > 0:                                          # :437
            latents, scale_range = weg_pre_step(denoise_fn, latents, i, t, enc, masks, focus_indices, weg, scale_range,
                                                len(scheduler.timesteps), weg_log)
        x = torch.cat([latents] * N_BRANCH)                                 # :499
        noise_pred, att_mats = denoise_fn(x, t, enc, masks)                 # :507
        att[int(t)] = [a.chunk(N_BRANCH)[-1] for a in att_mats]             # :519-523
        eps = guidance_combine(noise_pred, guidance_scale)                  # :527-541
        kw = {}
        if isinstance(scheduler, DDIMSchedulerOracle):
            kw["eta"] = eta                                                  # :426-429
        if step_noise is not None:
            kw["variance_noise"] = step_noise[i]
        latents = scheduler.step(eps, t, latents, **kw).prev_sample         # :544
        if record is not None:
            record.append(latents.clone())
    return latents.permute(1, 0, 2), att                                     # :548


# --------------------------------------------------------------------------- word-excitation guidance (WEG)
def weg_gaussian_kernel(kernel_size: int = 3, sigma: float = 0.5) -> Tensor:
    """operator/gaussian_smoothing.py:21-47 for dim=2, channels=1 (note the reference's exp(-((x - mean) / (2 std))^2))."""
    kernel = 1
    grids = torch.meshgrid([torch.arange(kernel_size, dtype=torch.float32)] * 2, indexing="ij")
    for mgrid in grids:
        mean = (kernel_size - 1) / 2
        kernel = kernel * (1 / (sigma * math.sqrt(2 * math.pi)) * torch.exp(-((mgrid - mean) / (2 * sigma)) ** 2))
    kernel = kernel / torch.sum(kernel)
    return kernel.view(1, 1, kernel_size, kernel_size)


def weg_max_attention(att_text: Tensor, focus_indices: Sequence[Sequence[int]], eot_indices: Tensor) -> List[List[Tensor]]:
    """tools/word_excitation_guidance.py:11-52 as the loops call it (smooth_attentions=True, normalize_eot=True):
    layer mean, drop BOS / EOS and beyond, softmax over the remaining tokens, 3x3 Gaussian smoothing with reflect padding,
    max over the 16 motion tokens for every focus token.  att_text [1, layers, 16, T]."""
    att = torch.mean(att_text, dim=1)                                       # :11-14
    assert att.shape[0] == 1, "EOS/BOS normalization only works for test batch size 1 currently"   # :25
    last_idx = int(eot_indices[0])                                          # :26
    a = torch.softmax(att[:, :, 1:last_idx], dim=-1)                        # :28-30
    a = F.conv2d(F.pad(a.unsqueeze(1), (1, 1, 1, 1), mode="reflect"), weg_gaussian_kernel().to(a)).squeeze(1)   # :33-36
    return [[a[b, :, i - 1].max(dim=-1)[0] for i in idxs] for b, idxs in enumerate(focus_indices)]   # :39-51


def weg_focus_loss(max_att: List[List[Tensor]]) -> Tensor:
    """tools/word_excitation_guidance.py:65-83 (no empty samples): mean_b mean_tokens max(0, 1 - max attention)."""
    losses = [torch.mean(torch.stack([torch.max(torch.zeros_like(t), 1.0 - t) for t in sample]), dim=-1) for sample in max_att]
    return torch.mean(torch.stack(losses, dim=-1))


def weg_evaluate(denoise_fn, latents: Tensor, t, enc_text, masks_text, focus_indices):
    """One WEG evaluation (convofusion.py:451-471 / :326-343): text-only branch forward, focus loss.
    Returns (loss, latents_with_grad)."""
    latents = latents.clone().detach().requires_grad_(True)
    _, att = denoise_fn(latents, t, enc_text, masks_text)
    eot = torch.argmax(masks_text["tlsn"].int(), dim=1) - 1                 # :463
    loss = weg_focus_loss(weg_max_attention(att[2], focus_indices, eot))    # :466-471
    return loss, latents


def weg_pre_step(denoise_fn, latents: Tensor, i: int, t, enc, masks, focus_indices, weg: dict, scale_range, n_steps: int,
                 log: Optional[list] = None):
    """convofusion.py:437-496 (== unbounded_synthesis.py:82-142): the latent update that precedes the guided step when
    focus tokens are given.  `scale_range` is re-assigned to the linspace array on every step exactly like the
    reference (:442-444), i.e. from the second step on its first two ELEMENTS are the end points.  Returns
    (latents, scale_range)."""
    scale_range = np.linspace(scale_range[0], scale_range[1], n_steps)      # :442-444
    enc_t = [e.chunk(N_BRANCH)[1] for e in enc]                             # :449
    masks_t = {k: (v.chunk(N_BRANCH)[1] if v is not None else v) for k, v in masks.items()}   # :450
    loss, latents = weg_evaluate(denoise_fn, latents, t, enc_t, masks_t, focus_indices)
    thresholds = weg["thresholds"]
    n_refine = 0
    if i in thresholds.keys() and loss > 1.0 - thresholds[i]:               # :474
        # iterative_refinement_step (:298-388)
        step_size = weg["scale_factor"] * np.sqrt(scale_range[i])
        target = max(0, 1.0 - thresholds[i])
        while loss > target:                                                # :326
            n_refine += 1
            loss, latents = weg_evaluate(denoise_fn, latents, t, enc_t, masks_t, focus_indices)
            if loss.all() != 0:                                             # :344
                grad = torch.autograd.grad(loss, [latents])[0]              # word_excitation_guidance.py:60
                latents = latents - step_size * grad
            if n_refine >= weg["max_refinement_steps"]:                     # :363
                break
        loss, latents = weg_evaluate(denoise_fn, latents, t, enc_t, masks_t, focus_indices)   # :368-387
    if i < weg["max_iter_to_alter"]:                                        # :490
        if loss.all() != 0:                                                 # :493
            grad = torch.autograd.grad(loss, [latents])[0]
            latents = latents - weg["scale_factor"] * np.sqrt(scale_range[i]) * grad        # :495
    if log is not None:
        log.append({"loss": float(loss.detach()), "n_refine": n_refine})
    return latents.detach(), scale_range


# unbounded_synthesis.py:82-87: the forecast loop hard-codes its WEG parameters ("TODO: move to config")
FORECAST_WEG = {"scale_factor": 100, "scale_range": (1.0, 0.5), "max_iter_to_alter": 800,
                "thresholds": {0: 0.05, 200: 0.4, 400: 0.6, 600: 0.8}, "max_refinement_steps": 300}


def diffusion_reverse_forecast(denoise_fn, scheduler, noise_scheduler, enc, masks, init_noise: Tensor,
                               num_steps: int, preseq: Optional[Tensor], guidance_scale: float = 7.5,
                               eta: float = 0.0, step_noise: Optional[Tensor] = None,
                               record: Optional[list] = None, focus_indices: Sequence[Sequence[int]] = (),
                               weg: Optional[dict] = None, weg_log: Optional[list] = None):
    """unbounded_synthesis.py:28-187; with focus_indices the latent update of :78-142 precedes every guided step (its
    parameters are the script's hard-coded ones, FORECAST_WEG, and `scale_range` is the fresh tuple on every step --
    unlike Convofusion._diffusion_reverse there is no array re-assignment here).  Reproduces the aliasing quirk:
    `latents = init_noise` (:66) shares storage, so the in-place inpaint at step 0 (:76) also
    rewrites init_noise[:, :preseq_len], which later steps then reuse as "noise" (:73)."""
    init_noise = init_noise.clone() * scheduler.init_noise_sigma            # :49
    scheduler.set_timesteps(num_steps)
    latents = init_noise                                                    # :66 (alias!)
    att_mats = None
    for i, t in enumerate(scheduler.timesteps):
        if preseq is not None:
            pl = preseq.shape[1]
            preseq_noise = init_noise.clone()                               # :73
            noised = noise_scheduler.add_noise(preseq.clone(), preseq_noise[:, :pl, :], t)  # :75
            latents[:, :pl, :] = noised                                     # :76
        if len(focus_indices) > 0:                                          # :78-142 (rebinds latents: the alias ends here)
            wp = weg if weg is not None else FORECAST_WEG
            latents, _ = weg_pre_step(denoise_fn, latents, i, t, enc, masks, focus_indices, wp, wp["scale_range"],
                                      len(scheduler.timesteps), weg_log)
        x = torch.cat([latents] * N_BRANCH)
        noise_pred, att_mats = denoise_fn(x, t, enc, masks)
        att_mats = [a.chunk(N_BRANCH)[-1] for a in att_mats]
        eps = guidance_combine(noise_pred, guidance_scale)
        kw = {}
        if isinstance(scheduler, DDIMSchedulerOracle):
            kw["eta"] = eta
        if step_noise is not None:
            kw["variance_noise"] = step_noise[i]
        latents = scheduler.step(eps, t, latents, **kw).prev_sample         # :181 (rebinds)
        if record is not None:
            record.append(latents.clone())
    return latents.permute(1, 0, 2), att_mats


def latents_to_vae_input(z: Tensor) -> Tensor:
    """convofusion.py:1027-1030: [16,B,d] -> [2,B,8,d] (token = 2*chunk + body/hand)."""
    ntok, bs, dim = z.shape
    return z.reshape(ntok // 2, 2, bs, dim).permute(1, 2, 0, 3)


def stitch_root(feats: Tensor, prev: Optional[Tensor]) -> Tensor:
    """unbounded_synthesis.py:461-465: re-anchor root x/z of this window on the previous one."""
    if prev is None:
        return feats
    feats = feats.clone()
    xz = torch.tensor([1.0, 0.0, 1.0], dtype=feats.dtype)
    feats[:, :, :3] = feats[:, :, :3] - feats[:, :1, :3] * xz
    feats[:, :, :3] = feats[:, :, :3] + prev[:, :1, :3] * xz
    return feats
