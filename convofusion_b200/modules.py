"""Host-side mirror of the reference's denoiser / VAE operator surface.

The classes here own `nn.Parameter`s under exactly the reference's names, so a reference checkpoint
loads with `load_state_dict(strict=True)`, and they keep the reference call signatures:

  Denoiser.forward(sample, timestep, encoder_hidden_states, lengths, mem_mask_dict)
      -> (eps [BG,16,128], 5 x att [BG,9,16,M_x])          convofusion/models/architectures/denoiser.py:173-386
  ConvoFusionVae.decode(z [2,B,8,128], lengths) -> [B,T,189]  convofusion/models/architectures/vae.py:268-372

All arithmetic runs in libconvofusion_b200.so through the C ABI; nothing here computes on the CPU and
nothing falls back to torch ops: a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib
from .pack import pack_denoiser, pack_vae

STREAMS = ("spkemb", "alsn", "tlsn", "apb", "lsnemb")   # cross_attention.py:579


def _sine_pe(d_model: int, max_len: int = 1024) -> Tensor:
    # position_encoding.py:119-127 buffer 'pe' [max_len, 1, d_model]
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(1)


class _PE(nn.Module):
    def __init__(self, d_model):
        super().__init__()
        self.register_buffer("pe", _sine_pe(d_model))


class _MHA(nn.Module):
    """Parameter holder with nn.MultiheadAttention's names (in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)


class _TimeBlock(nn.Module):
    # cross_attention.py:411-424: emb_layers = (SiLU, Linear), out_layers = (SiLU, Dropout, Linear)
    def __init__(self, d):
        super().__init__()
        self.emb_layers = nn.Sequential(nn.Identity(), nn.Linear(d, 2 * d))
        self.norm = nn.LayerNorm(d)
        self.out_layers = nn.Sequential(nn.Identity(), nn.Identity(), nn.Linear(d, d))


class _DenoiserLayer(nn.Module):
    # cross_attention.py:444-489
    def __init__(self, d, ff):
        super().__init__()
        self.self_attn = _MHA(d)
        self.time_block1 = _TimeBlock(d)
        self.multihead_attn_spkemb = _MHA(d)
        self.multihead_attn_tlsn = _MHA(d)
        self.multihead_attn_alsn = _MHA(d)
        self.multihead_attn_apb = _MHA(d)
        self.multihead_attn_lsnemb = _MHA(d)
        self.att_fuser = nn.Linear(5 * d, d)
        self.time_block2 = _TimeBlock(d)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        for n in ("norm1", "norm2", "norm3", "spkemb_norm", "alsn_norm", "tlsn_norm", "apb_norm", "lsnemb_norm"):
            setattr(self, n, nn.LayerNorm(d))


class _DecoderStack(nn.Module):
    def __init__(self, d, ff, n_layers):
        super().__init__()
        self.layers = nn.ModuleList([_DenoiserLayer(d, ff) for _ in range(n_layers)])
        self.norm = nn.LayerNorm(d)


class _TimestepEmbedding(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.linear_1 = nn.Linear(d, d)
        self.linear_2 = nn.Linear(d, d)


def _precision_id(p: str) -> int:
    if p not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {p!r}")
    return _lib.F32 if p == "fp32" else _lib.BF16


_lane_state = threading.local()
_lane_lock = threading.RLock()   # handle creation (module-level: nn.Modules must stay deep-copyable)


@contextlib.contextmanager
def lane(index: int, chains: Optional[int] = None):
    """Selects the device handle the calling thread's Denoiser / ConvoFusionVae calls run on.

    A handle owns one workspace and one captured step graph, so it serves one call at a time.  Lane k > 0 is an
    extra handle over the SAME packed weights, created on first use; `SamplerPool` (pool.py) runs one host thread +
    one stream per lane to keep several independent batches in flight on one GPU."""
    prev, prev_chains = getattr(_lane_state, "index", 0), getattr(_lane_state, "chains", None)
    _lane_state.index, _lane_state.chains = int(index), chains
    try:
        yield
    finally:
        _lane_state.index, _lane_state.chains = prev, prev_chains


def current_lane() -> int:
    return getattr(_lane_state, "index", 0)


def current_chains() -> Optional[int]:
    """Concurrent chains per captured step requested for the calling thread's lane (None = the module's setting)."""
    return getattr(_lane_state, "chains", None)


class _CudaModule(nn.Module):
    """Shared handle management: weights are (re)packed lazily whenever parameters may have changed -- `.to()` /
    `load_state_dict` / `set_precision` invalidate eagerly, and every call compares a key built from each parameter's
    storage pointer and in-place version counter (`p.data.copy_`, `optimizer.step`, `torch.nn.init`, EMA swaps), so
    stale folded weights / captured graphs are never used silently."""

    def __init__(self):
        super().__init__()
        self._handle = None      # lane 0
        self._lanes = {}         # lane k > 0 -> handle over the same packed weights
        self._packed = None      # keeps packed device tensors + ctypes structs alive
        self._pack_key = None
        self.precision = os.environ.get("CONVOFUSION_B200_PRECISION", "bf16")

    def _destroy(self, handle):
        raise NotImplementedError

    def _create(self, packed):
        raise NotImplementedError

    def _invalidate(self):
        if self._handle is not None:
            self._destroy(self._handle)
        for h in getattr(self, "_lanes", {}).values():
            self._destroy(h)
        self._handle, self._packed, self._pack_key = None, None, None
        self._lanes = {}

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure(self):
        """Handle of the calling thread's lane (packs the weights / creates the handle on first use)."""
        k = current_lane()
        if self._handle is not None and self._pack_key != self._param_key():
            if k != 0:
                raise _lib.CfbError("parameters were modified in place while lanes are in flight; call pack() from "
                                    "the main thread before SamplerPool.map")
            self._invalidate()                       # parameters were edited in place since pack()
        if k == 0 and self._handle is not None:
            return self._handle
        with _lane_lock:
            if self._handle is None:
                self.pack()
                self._pack_key = self._param_key()
            if k == 0:
                return self._handle
            if k not in self._lanes:
                with torch.cuda.device(self._device()):
                    self._lanes[k] = self._create(self._packed)
            return self._lanes[k]

    def __getstate__(self):
        # copy.deepcopy / pickle: device handles and packed weights belong to THIS object (a copied pointer would be
        # destroyed twice); the copy re-packs lazily on its first call
        state = self.__dict__.copy()
        state["_handle"], state["_lanes"], state["_packed"], state["_pack_key"] = None, {}, None, None
        return state

    def _apply(self, fn, *a, **k):   # .to() / .cuda() / .float()
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    def set_precision(self, precision: str):
        _precision_id(precision)
        if precision != self.precision:
            self._invalidate()
            self.precision = precision
        return self

    def _device(self) -> torch.device:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.CfbError(f"{type(self).__name__} parameters live on {dev}: convofusion_b200 runs on a B200 "
                                "only and has no CPU path; call .cuda() first")
        return dev

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass


class Denoiser(_CudaModule):
    """Drop-in for convofusion.models.architectures.denoiser.Denoiser (arch=trans_dec, text+audio)."""

    def __init__(self, ablation=None, nfeats: int = 263, condition: str = "text", latent_dim: list = [1, 256],
                 ff_size: int = 1024, num_layers: int = 6, num_heads: int = 4, dropout: float = 0.1,
                 normalize_before: bool = False, activation: str = "gelu", flip_sin_to_cos: bool = True,
                 return_intermediate_dec: bool = False, position_embedding: str = "learned",
                 arch: str = "trans_enc", freq_shift: int = 0, guidance_scale: float = 7.5,
                 guidance_uncondp: float = 0.1, text_encoded_dim: int = 768, audio_encoded_dim: int = 512,
                 nclasses: int = 10, precision: Optional[str] = None, **kwargs) -> None:
        super().__init__()
        # denoiser.py:78,153,115: only the shipped configuration is implemented
        if condition not in ("text+audio", "textaudio_uncond"):
            raise TypeError(f"condition type {condition} not supported")
        if arch != "trans_dec":
            raise ValueError(f"Not supported architechure{arch}!")
        if ablation is not None and getattr(ablation, "DIFF_PE_TYPE", "convofusion") != "convofusion":
            raise ValueError("Not Support PE type")
        if ablation is not None and getattr(ablation, "VAE_TYPE", "convofusion") == "no":
            raise ValueError("diffusion-only (VAE_TYPE='no') is outside the B200 hot path")
        if not normalize_before or activation != "gelu" or not flip_sin_to_cos or freq_shift != 0:
            raise ValueError("only normalize_before=True, activation='gelu', flip_sin_to_cos=True, freq_shift=0 "
                             "(configs/modules/denoiser.yaml) is implemented")
        if position_embedding not in ("sine", "v2"):
            raise ValueError(f"not supported {position_embedding}")
        if ablation is not None and getattr(ablation, "CAUSAL_ATTN", False):
            raise ValueError("CAUSAL_ATTN=True is not implemented")
        d = text_encoded_dim
        self.latent_dim = latent_dim[-1]
        self.text_encoded_dim = d
        self.audio_encoded_dim = audio_encoded_dim
        self.condition = condition
        self.arch = arch
        self.num_layers, self.num_heads, self.ff_size = num_layers, num_heads, ff_size
        self.n_tokens = 16    # 8 chunks x (body, hand): convofusion.py:412-413
        self.latent_embd = nn.Linear(self.latent_dim, d)
        self.latent_proj = nn.Linear(d, self.latent_dim)
        self.time_embedding = _TimestepEmbedding(d)
        self.query_pos = _PE(d)
        self.mem_pos = _PE(d)
        self.bh_embedding = nn.Embedding(2, d)
        self.condition_embedding = nn.Embedding(5, d)
        self.cond_params = nn.Parameter(1 / 5 * torch.ones(5))
        self.decoder = _DecoderStack(d, ff_size, num_layers)
        self.step_chains = 0     # concurrent chains of the captured sampling step (0 = library default)
        if precision is not None:
            self.set_precision(precision)

    # ---- handle
    def _destroy(self, handle):
        _lib.lib().cfb_denoiser_destroy(handle)

    def _create(self, packed):
        h = C.c_void_p()
        _lib.check(_lib.lib().cfb_denoiser_create(C.byref(packed["struct"]), C.byref(h)))
        if "struct16" in packed:      # 16-bit handles: fp16 matrices packed from the fp32 parameters (pack.py)
            try:
                _lib.check(_lib.lib().cfb_denoiser_attach_f16_weights(h, C.byref(packed["struct16"])))
            except Exception:
                _lib.lib().cfb_denoiser_destroy(h)
                raise
        return h

    def pack(self):
        """(Re)build the packed weights and the device handle.  Called lazily by forward()/sample()."""
        dev = self._device()
        self._invalidate()
        packed = pack_denoiser(self.state_dict(), "", self.num_layers, self.num_heads, self.n_tokens,
                               _precision_id(self.precision), dev)
        with torch.cuda.device(dev):
            # the casts / copies above ran on the caller's current stream; other streams (SamplerPool lanes) read the
            # packed weights too, so they are made globally visible before the handle is published (one-time cost)
            torch.cuda.current_stream(dev).synchronize()
            h = self._create(packed)
        self._handle, self._packed = h, packed
        self._pack_key = self._param_key()
        return self

    # ---- memory description
    @staticmethod
    def _memory(enc: Sequence[Tensor], masks: Dict[str, Optional[Tensor]], slots: Optional[Sequence[Optional[Tensor]]],
                keep: list, slots_host: Optional[Sequence[Tensor]] = None) -> "_lib.Memory":
        mem = _lib.Memory()
        for x, (name, e) in enumerate(zip(STREAMS, enc)):
            if e.dim() != 3:
                raise ValueError(f"encoder_hidden_states[{x}] must be [slots, len, d], got {tuple(e.shape)}")
            e = e.detach().to(torch.float32).contiguous()
            keep.append(e)
            mem.cond[x] = e.data_ptr()
            mem.n_slots[x], mem.len[x] = e.shape[0], e.shape[1]
            m = masks.get(name) if masks else None
            if m is not None:
                if tuple(m.shape) != (e.shape[0], e.shape[1]):
                    raise ValueError(f"mem_mask_dict[{name!r}] has shape {tuple(m.shape)}, expected {tuple(e.shape[:2])}")
                m8 = m.to(device=e.device, dtype=torch.uint8).contiguous()
                keep.append(m8)
                mem.mask[x] = m8.data_ptr()
            s = slots[x] if slots is not None else None
            if s is not None:
                s32 = s.to(device=e.device, dtype=torch.int32).contiguous()
                keep.append(s32)
                mem.slot[x] = s32.data_ptr()
                if slots_host is not None:
                    sh = slots_host[x].to(device="cpu", dtype=torch.int32).contiguous()
                    if sh.numel() != s32.numel():
                        raise ValueError("slots_host does not match slots")
                    keep.append(sh)
                    mem.slot_host[x] = sh.data_ptr()
        return mem

    # ---- reference surface
    def forward(self, sample: Tensor, timestep, encoder_hidden_states, lengths=None,
                mem_mask_dict: Optional[dict] = None, return_attention: bool = True, **kwargs):
        dev = self._device()
        if sample.device != dev:
            raise _lib.CfbError(f"sample is on {sample.device}, model on {dev}")
        h = self._ensure()
        bg, ntok, lat = sample.shape
        if ntok != self.n_tokens or lat != self.latent_dim:
            raise ValueError(f"sample must be [B,{self.n_tokens},{self.latent_dim}], got {tuple(sample.shape)}")
        if len(encoder_hidden_states) != 5:
            raise ValueError("encoder_hidden_states must be (spk_emb, alsn, tlsn, apb, lsnemb)")
        keep: list = []
        x = sample.detach().to(torch.float32).contiguous()
        mem = self._memory(encoder_hidden_states, mem_mask_dict or {}, None, keep)
        for i in range(5):
            if mem.n_slots[i] != bg:
                raise ValueError(f"encoder_hidden_states[{i}] has batch {mem.n_slots[i]}, sample has {bg}")
        eps = torch.empty_like(x)
        att = [torch.empty(bg, self.num_layers, ntok, mem.len[i], device=dev, dtype=torch.float32)
               for i in range(5)] if return_attention else None
        att_ptrs = (C.c_void_p * 5)(*[a.data_ptr() for a in att]) if att is not None else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cfb_denoiser_forward(h, x.data_ptr(), bg, int(timestep), C.byref(mem), eps.data_ptr(),
                                                       att_ptrs, _lib.stream_ptr()))
        return (eps, att)

    # ---- word-excitation guidance: attention maps of one stream and the gradient of a loss on them w.r.t. the latents
    def weg_forward(self, sample: Tensor, timestep, encoder_hidden_states, mem_mask_dict: Optional[dict] = None,
                    stream: int = 2) -> Tensor:
        """Denoiser.forward on the text-only branch as the reference's WEG calls it (convofusion.py:455-461), keeping
        the activations on the device; returns the attention maps of `stream` (2 = listener text, what the loss reads:
        :466) as [B, layers, 16, M].  fp32 handles only.  Follow with `weg_backward`."""
        if self.precision != "fp32":
            raise _lib.CfbError("word-excitation guidance differentiates the denoiser in fp32: use a precision='fp32' Denoiser")
        dev = self._device()
        h = self._ensure()
        bg, ntok, lat = sample.shape
        if ntok != self.n_tokens or lat != self.latent_dim:
            raise ValueError(f"sample must be [B,{self.n_tokens},{self.latent_dim}], got {tuple(sample.shape)}")
        keep: list = []
        x = sample.detach().to(device=dev, dtype=torch.float32).contiguous()
        mem = self._memory(encoder_hidden_states, mem_mask_dict or {}, None, keep)
        for i in range(5):
            if mem.n_slots[i] != bg:
                raise ValueError(f"encoder_hidden_states[{i}] has batch {mem.n_slots[i]}, sample has {bg}")
        att = torch.empty(bg, self.num_layers, ntok, mem.len[stream], device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cfb_denoiser_weg_forward(h, x.data_ptr(), bg, int(timestep), C.byref(mem), int(stream),
                                                           att.data_ptr(), _lib.stream_ptr()))
        self._weg_shape = (tuple(att.shape), (bg, ntok, lat))
        return att

    def weg_backward(self, d_att: Tensor) -> Tensor:
        """dLoss/dsample [B,16,latent] for dLoss/dAtt of the last `weg_forward` (torch.autograd.grad(loss, latents) of
        tools/word_excitation_guidance.py:60, through hand-written input-gradient kernels, csrc/weg.cu)."""
        att_shape, x_shape = getattr(self, "_weg_shape", (None, None))
        if att_shape is None or tuple(d_att.shape) != att_shape:
            raise ValueError(f"d_att must have the shape of the last weg_forward's maps {att_shape}")
        dev = self._device()
        h = self._ensure()
        g = d_att.detach().to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty(x_shape, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cfb_denoiser_weg_backward(h, g.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
        return out

    def sample(self, scheduler, enc: Sequence[Tensor], masks: Dict[str, Optional[Tensor]],
               slots: Sequence[Optional[Tensor]], latents: Tensor, num_steps: int, guidance_scale: float = 7.5,
               eta: float = 0.0, n_branch: int = 7, full_last: Optional[bool] = None,
               slots_host: Optional[Sequence[Tensor]] = None, step_noise: Optional[Tensor] = None,
               preseq: Optional[Tensor] = None, noise_scheduler=None, record: bool = False,
               return_attention: bool = False, use_graph: bool = True):
        """Whole guided reverse loop on the device (convofusion.py:391-549 / unbounded_synthesis.py:28-187).

        enc/masks/slots describe de-duplicated conditioning memory: enc[x] is [n_slots_x, M_x, 512] and
        slots[x] [n_branch*B] picks the slot every (branch, clip) attends to (`slots_host`: the same tables on the
        host, so the library need not read them back); `full_last`: the last evaluated branch is the weight-0
        full-cond one (default: n_branch == 7).  `latents` [B,16,128] is the
        initial noise already scaled by init_noise_sigma.  Returns (latents [B,16,128], record or None,
        attention maps or None)."""
        dev = self._device()
        h = self._ensure()
        B = latents.shape[0]
        table = scheduler.step_table(num_steps, eta=eta, noise_scheduler=noise_scheduler)
        keep: list = []
        mem = self._memory(enc, masks, slots, keep, slots_host)
        if full_last is None:
            full_last = n_branch == 7
        x = latents.detach().to(device=dev, dtype=torch.float32).contiguous().clone()
        sched = _lib.Schedule()
        ts = np.ascontiguousarray(table["timesteps"], dtype=np.int64)
        cf = np.ascontiguousarray(table["coef"], dtype=np.float32)
        sched.kind, sched.n_steps, sched.clip_sample = table["kind"], len(ts), int(table["clip_sample"])
        sched.guidance_scale = float(guidance_scale)
        sched.timesteps = ts.ctypes.data_as(C.POINTER(C.c_int64))
        sched.coef = cf.ctypes.data_as(C.POINTER(C.c_float))
        if table["needs_noise"] and step_noise is None:
            raise ValueError("this schedule draws noise every step (DDPM / eta>0): pass step_noise [steps,B,16,128]")
        if step_noise is not None:
            step_noise = step_noise.to(device=dev, dtype=torch.float32).contiguous()
            if tuple(step_noise.shape) != (len(ts), B, self.n_tokens, self.latent_dim):
                raise ValueError(f"step_noise must be {(len(ts), B, self.n_tokens, self.latent_dim)}")
        pl = 0
        if preseq is not None:
            preseq = preseq.to(device=dev, dtype=torch.float32).contiguous()
            pl = preseq.shape[1]
        rec = torch.empty(len(ts), B, self.n_tokens, self.latent_dim, device=dev) if record else None
        att, att_ptrs = None, None
        if return_attention:
            att = [torch.empty(len(ts), B, self.num_layers, self.n_tokens, mem.len[i], device=dev) for i in range(5)]
            att_ptrs = (C.c_void_p * 5)(*[a.data_ptr() for a in att])
        with torch.cuda.device(dev):
            chains = current_chains()
            _lib.check(_lib.lib().cfb_denoiser_set_chains(h, int(self.step_chains if chains is None else chains)))
            _lib.check(_lib.lib().cfb_sample(h, C.byref(sched), C.byref(mem), B, n_branch, int(bool(full_last)), x.data_ptr(),
                                             _lib.ptr(step_noise), _lib.ptr(preseq), pl, _lib.ptr(rec), att_ptrs,
                                             int(use_graph), _lib.stream_ptr()))
        return x, rec, att


# ------------------------------------------------------------------------------------- VAE
class _EncoderLayer(nn.Module):   # cross_attention.py:250-268 (VAE encode side: parameters only)
    def __init__(self, d, ff):
        super().__init__()
        self.self_attn = _MHA(d)
        self.linear1, self.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
        self.norm1, self.norm2 = nn.LayerNorm(d), nn.LayerNorm(d)


class _VaeDecoderLayer(nn.Module):   # cross_attention.py:311-332
    def __init__(self, d, ff):
        super().__init__()
        self.self_attn = _MHA(d)
        self.multihead_attn = _MHA(d)
        self.linear1, self.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)


class _Skip(nn.Module):   # cross_attention.py:18-39 / 66-87
    def __init__(self, layer_cls, d, ff, n_layers):
        super().__init__()
        nb = (n_layers - 1) // 2
        self.input_blocks = nn.ModuleList([layer_cls(d, ff) for _ in range(nb)])
        self.middle_block = layer_cls(d, ff)
        self.output_blocks = nn.ModuleList([layer_cls(d, ff) for _ in range(nb)])
        self.linear_blocks = nn.ModuleList([nn.Linear(2 * d, d) for _ in range(nb)])
        self.norm = nn.LayerNorm(d)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class ConvoFusionVae(_CudaModule):
    """Drop-in for convofusion.models.architectures.vae.ConvoFusionVae (encode and decode on the B200).
    The class name is load-bearing: convofusion.py:67-71 derives vae_type from it."""

    def __init__(self, ablation=None, nfeats: int = 189, latent_dim: list = [1, 256], ff_size: int = 1024,
                 num_layers: int = 9, num_heads: int = 4, dropout: float = 0.1, arch: str = "all_encoder",
                 normalize_before: bool = False, activation: str = "gelu", position_embedding: str = "learned",
                 precision: Optional[str] = None, **kwargs) -> None:
        super().__init__()
        if arch != "encoder_decoder":
            raise ValueError("Not support architecture!")
        if ablation is not None and getattr(ablation, "PE_TYPE", "convofusion") != "convofusion":
            raise ValueError("Not support position encoding type!")
        if ablation is not None and getattr(ablation, "MLP_DIST", False):
            raise ValueError("MLP_DIST=True is not implemented")
        if not normalize_before or activation != "gelu" or position_embedding not in ("sine", "v2"):
            raise ValueError("only normalize_before=True, activation='gelu', position_embedding='sine' "
                             "(configs/modules/motion_vae.yaml) is implemented")
        self.latent_size, self.latent_dim = latent_dim[0], latent_dim[-1]
        self.body_nfeats, self.hands_nfeats = 23 * 3, 40 * 3
        self.arch, self.num_layers, self.num_heads, self.ff_size = arch, num_layers, num_heads, ff_size
        d = self.latent_dim
        self.query_pos_encoder, self.query_pos_decoder, self.mem_pos_decoder = _PE(d), _PE(d), _PE(d)
        self.body_encoder = _Skip(_EncoderLayer, d, ff_size, num_layers)
        self.hands_encoder = _Skip(_EncoderLayer, d, ff_size, num_layers)
        self.body_decoder = _Skip(_VaeDecoderLayer, d, ff_size, num_layers)
        self.hands_decoder = _Skip(_VaeDecoderLayer, d, ff_size, num_layers)
        self.body_global_motion_token = nn.Parameter(torch.randn(self.latent_size * 2, d))
        self.hands_global_motion_token = nn.Parameter(torch.randn(self.latent_size * 2, d))
        self.body_skel_embedding = nn.Linear(self.body_nfeats, d)
        self.hands_skel_embedding = nn.Linear(self.hands_nfeats, d)
        self.body_final_layer = nn.Linear(d, self.body_nfeats)
        self.hands_final_layer = nn.Linear(d, self.hands_nfeats)
        if precision is not None:
            self.set_precision(precision)

    def _destroy(self, handle):
        _lib.lib().cfb_vae_destroy(handle)

    def _create(self, packed):
        h = C.c_void_p()
        _lib.check(_lib.lib().cfb_vae_create(C.byref(packed["struct"]), C.byref(h)))
        return h

    def pack(self):
        dev = self._device()
        self._invalidate()
        packed = pack_vae(self.state_dict(), "", self.num_layers, self.num_heads, self.ff_size,
                          _precision_id(self.precision), dev)
        with torch.cuda.device(dev):
            # the casts / copies above ran on the caller's current stream; other streams (SamplerPool lanes) read the
            # packed weights too, so they are made globally visible before the handle is published (one-time cost)
            torch.cuda.current_stream(dev).synchronize()
            h = self._create(packed)
        self._handle, self._packed = h, packed
        self._pack_key = self._param_key()
        return self

    def decode(self, z: Tensor, lengths: List[int]) -> Tensor:
        dev = self._device()
        handle = self._ensure()
        if z.dim() != 4 or z.shape[0] != 2 or z.shape[-1] != self.latent_dim:
            raise ValueError(f"z must be [2, B, n_chunks, {self.latent_dim}], got {tuple(z.shape)}")
        _, bs, n_chunks, _ = z.shape
        lengths = [int(l) for l in lengths]
        if len(lengths) != bs:
            raise ValueError(f"{len(lengths)} lengths for batch {bs}")
        nframes = max(lengths)   # lengths_to_mask: max_len = max(lengths) (temos_utils.py:15)
        zc = z.detach().to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty(bs, nframes, self.body_nfeats + self.hands_nfeats, device=dev, dtype=torch.float32)
        lens = (C.c_int32 * bs)(*lengths)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cfb_vae_decode(handle, zc.data_ptr(), bs, n_chunks, nframes, lens,
                                                 out.data_ptr(), _lib.stream_ptr()))
        return out

    def encode_params(self, features: Tensor, lengths: Optional[List[int]] = None) -> Tuple[Tensor, Tensor, Tensor]:
        """Deterministic part of encode (vae.py:162-260): (mu, std [2, B*T/16, d], root-subtracted features)."""
        dev = self._device()
        handle = self._ensure()
        if features.dim() != 3 or features.shape[-1] != self.body_nfeats + self.hands_nfeats:
            raise ValueError(f"features must be [B, T, {self.body_nfeats + self.hands_nfeats}], got {tuple(features.shape)}")
        bs, nframes, nfeats = features.shape
        if lengths is None:
            lengths = [len(f) for f in features]
        lengths = [int(l) for l in lengths]
        if len(lengths) != bs:
            raise ValueError(f"{len(lengths)} lengths for batch {bs}")
        if nframes % 16 or max(lengths) != nframes:
            # the reference reshapes lengths_to_mask(lengths) ([B, max(lengths)]) to [B*T/16, 16] (vae.py:173,187)
            raise ValueError("encode needs T to be a multiple of 16 and max(lengths) == T")
        x = features.detach().to(device=dev, dtype=torch.float32).contiguous()
        n = bs * (nframes // 16)
        mu = torch.empty(2 * self.latent_size, n, self.latent_dim, device=dev, dtype=torch.float32)
        std = torch.empty_like(mu)
        feats = torch.empty_like(x)
        lens = (C.c_int32 * bs)(*lengths)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cfb_vae_encode(handle, x.data_ptr(), bs, nframes, lens, mu.data_ptr(),
                                                 std.data_ptr(), feats.data_ptr(), _lib.stream_ptr()))
        return mu, std, feats

    def encode(self, features: Tensor, lengths: Optional[List[int]] = None):
        """vae.py:162-266: returns (latent [2, B, T/16, d], Normal(mu, std), root-subtracted features)."""
        if self.latent_size != 1:
            raise ValueError("latent_size != 1 is not implemented")
        mu, std, feats = self.encode_params(features, lengths)
        dist = torch.distributions.Normal(mu, std)
        latent = dist.rsample()                                   # vae.py:261-262 (torch RNG, like the reference)
        bs = features.shape[0]
        return latent.reshape(-1, bs, features.shape[1] // 16, self.latent_dim), dist, feats

    def forward(self, features: Tensor, lengths: Optional[List[int]] = None):
        """vae.py:145-160: encode, then decode the sampled latent."""
        z, dist, _ = self.encode(features, lengths)
        if lengths is None:
            lengths = [len(f) for f in features]
        return self.decode(z, lengths), z, dist
