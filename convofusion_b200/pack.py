"""Weight packing: reference state_dict layout -> the kernel-facing layout of include/convofusion_b200.h.

Done once per load (host, float64): matrices go to the handle's precision, vectors stay float32.

Folded cross-attention (exact algebra): for stream x of layer l with
  (Wq,bq | Wk,bk | Wv,bv) = multihead_attn_x.in_proj, (O,bo) = out_proj, (g,b) = {x}_norm, F_x = att_fuser[:, x-block]
and xhat = LayerNorm-without-affine(memory):
  scores = softmax_j( (A_x n + a_x) . xhat_j ),  A_x = (Wk diag g)^T Wq / sqrt(d),  a_x = (Wk diag g)^T bq / sqrt(d)
      (the dropped term q.(Wk b + bk) is constant over keys j, softmax is shift invariant)
  fuser(cat_x out_proj_x(P_x V_x)) = sum_x G_x (P_x xhat) + c,  G_x = F_x O Wv diag g,
      c = att_fuser.bias + sum_x F_x (O (Wv b + bv) + bo)      (rows of P_x sum to 1)
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict

import torch
from torch import Tensor

from . import _lib

STREAMS = ("spkemb", "alsn", "tlsn", "apb", "lsnemb")


class _Packer:
    def __init__(self, device, precision, f16: bool = False):
        self.device, self.precision, self.keep, self.f16 = device, precision, [], f16

    def mat(self, t: Tensor) -> int:
        # 16-bit handles: bf16, or fp16 where the handle's kernels take fp16 operands (the VAE, cfb_get_vae_f16)
        dt = (torch.float16 if self.f16 else torch.bfloat16) if self.precision == _lib.BF16 else torch.float32
        t = t.detach().to(dtype=dt).contiguous().to(self.device)
        self.keep.append(t)
        return t.data_ptr()

    def vec(self, t: Tensor) -> int:
        t = t.detach().to(dtype=torch.float32).contiguous().to(self.device)
        self.keep.append(t)
        return t.data_ptr()


def fold_cross_attention(sd: Dict[str, Tensor], lp: str, d: int):
    """Returns (w_qx [5d,d], b_qx [5d], w_fu [d,5d], b_fu [d]) in float64."""
    f64 = lambda k: sd[k].detach().double().cpu()
    Fw, Fb = f64(lp + "att_fuser.weight"), f64(lp + "att_fuser.bias")
    inv = 1.0 / math.sqrt(d)    # single head: head_dim = d
    A, a, G = [], [], []
    c = Fb.clone()
    for x, name in enumerate(STREAMS):
        W, b = f64(lp + f"multihead_attn_{name}.in_proj_weight"), f64(lp + f"multihead_attn_{name}.in_proj_bias")
        O, bo = f64(lp + f"multihead_attn_{name}.out_proj.weight"), f64(lp + f"multihead_attn_{name}.out_proj.bias")
        g, be = f64(lp + f"{name}_norm.weight"), f64(lp + f"{name}_norm.bias")
        Wq, Wk, Wv = W[:d], W[d:2 * d], W[2 * d:]
        bq, bv = b[:d], b[2 * d:]
        Wk_g, Wv_g = Wk * g[None, :], Wv * g[None, :]
        A.append(Wk_g.T @ Wq * inv)
        a.append(Wk_g.T @ bq * inv)
        Fx = Fw[:, x * d:(x + 1) * d]
        G.append(Fx @ O @ Wv_g)
        c = c + Fx @ (O @ (Wv @ be + bv) + bo)
    return torch.cat(A, 0), torch.cat(a, 0), torch.cat(G, 1), c


def pack_denoiser(sd: Dict[str, Tensor], prefix: str, n_layers: int, n_heads: int, n_tokens: int, precision: int,
                  device) -> dict:
    p = prefix
    pk = _Packer(device, precision)
    # 16-bit handles: every matrix is packed twice from the fp32 state_dict -- bf16 (the handle's own weights) and fp16
    # (cfb_denoiser_attach_f16_weights: the operand of the fp16 x fp16 products, 11 instead of 8 significant bits)
    pk16 = _Packer(device, precision, f16=True) if precision == _lib.BF16 else None
    layers16 = (_lib.DenoiserLayer * n_layers)() if pk16 else None

    def mat2(t, obj16, field):
        if pk16 is not None:
            setattr(obj16, field, pk16.mat(t))
        return pk.mat(t)
    d, lat = sd[p + "latent_embd.weight"].shape
    ff = sd[p + "decoder.layers.0.linear1.weight"].shape[0]
    layers = (_lib.DenoiserLayer * n_layers)()
    tbw, tbb = [], []
    zx, az, yx = [[] for _ in STREAMS], [[] for _ in STREAMS], [[] for _ in STREAMS]
    for l in range(n_layers):
        lp = f"{p}decoder.layers.{l}."
        L = layers[l]
        L16 = layers16[l] if pk16 else None
        L.ln1_g, L.ln1_b = pk.vec(sd[lp + "norm1.weight"]), pk.vec(sd[lp + "norm1.bias"])
        L.w_in, L.b_in = mat2(sd[lp + "self_attn.in_proj_weight"], L16, "w_in"), pk.vec(sd[lp + "self_attn.in_proj_bias"])
        L.w_so, L.b_so = mat2(sd[lp + "self_attn.out_proj.weight"], L16, "w_so"), pk.vec(sd[lp + "self_attn.out_proj.bias"])
        L.tb1_g, L.tb1_b = pk.vec(sd[lp + "time_block1.norm.weight"]), pk.vec(sd[lp + "time_block1.norm.bias"])
        L.w_tb1 = mat2(sd[lp + "time_block1.out_layers.2.weight"], L16, "w_tb1")
        L.b_tb1 = pk.vec(sd[lp + "time_block1.out_layers.2.bias"])
        L.ln2_g, L.ln2_b = pk.vec(sd[lp + "norm2.weight"]), pk.vec(sd[lp + "norm2.bias"])
        w_qx, b_qx, w_fu, b_fu = fold_cross_attention(sd, lp, d)
        L.w_qx, L.b_qx, L.w_fu, L.b_fu = mat2(w_qx, L16, "w_qx"), pk.vec(b_qx), mat2(w_fu, L16, "w_fu"), pk.vec(b_fu)
        for x in range(len(STREAMS)):
            zx[x].append(w_qx[x * d:(x + 1) * d].T)          # A_x^T: memory row -> key in query space
            az[x].append(b_qx[x * d:(x + 1) * d][None, :])
            yx[x].append(w_fu[:, x * d:(x + 1) * d])         # G_x: memory row -> its residual contribution
        L.tb2_g, L.tb2_b = pk.vec(sd[lp + "time_block2.norm.weight"]), pk.vec(sd[lp + "time_block2.norm.bias"])
        L.w_tb2 = mat2(sd[lp + "time_block2.out_layers.2.weight"], L16, "w_tb2")
        L.b_tb2 = pk.vec(sd[lp + "time_block2.out_layers.2.bias"])
        L.ln3_g, L.ln3_b = pk.vec(sd[lp + "norm3.weight"]), pk.vec(sd[lp + "norm3.bias"])
        L.w_ff1, L.b_ff1 = mat2(sd[lp + "linear1.weight"], L16, "w_ff1"), pk.vec(sd[lp + "linear1.bias"])
        L.w_ff2, L.b_ff2 = mat2(sd[lp + "linear2.weight"], L16, "w_ff2"), pk.vec(sd[lp + "linear2.bias"])
        for tb in ("time_block1", "time_block2"):
            tbw.append(sd[lp + f"{tb}.emb_layers.1.weight"].detach().float().cpu())
            tbb.append(sd[lp + f"{tb}.emb_layers.1.bias"].detach().float().cpu())
    w = _lib.DenoiserWeights()
    w16 = _lib.DenoiserWeights() if pk16 else None
    pe_q = sd[p + "query_pos.pe"].detach().float().cpu()[:, 0]
    pe_m = sd[p + "mem_pos.pe"].detach().float().cpu()[:, 0]
    bh = sd[p + "bh_embedding.weight"].detach().float().cpu()
    tok = torch.arange(n_tokens)
    # latent_embd.bias + bh_embedding[tok % 2] + query_pos.pe[tok // 2]   (denoiser.py:187,316-326)
    tok_bias = sd[p + "latent_embd.bias"].detach().float().cpu()[None, :] + bh[tok % 2] + pe_q[tok // 2]
    w.d_model, w.latent_dim, w.n_tokens, w.n_layers, w.n_heads, w.ff_size = d, lat, n_tokens, n_layers, n_heads, ff
    w.precision, w.pe_len = precision, pe_m.shape[0]
    w.w_embed, w.tok_bias = mat2(sd[p + "latent_embd.weight"], w16, "w_embed"), pk.vec(tok_bias)
    w.w_t1, w.b_t1 = pk.vec(sd[p + "time_embedding.linear_1.weight"]), pk.vec(sd[p + "time_embedding.linear_1.bias"])
    w.w_t2, w.b_t2 = pk.vec(sd[p + "time_embedding.linear_2.weight"]), pk.vec(sd[p + "time_embedding.linear_2.bias"])
    w.w_tbmod, w.b_tbmod = pk.vec(torch.cat(tbw, 0)), pk.vec(torch.cat(tbb, 0))
    w.stream_emb, w.pe_mem = pk.vec(sd[p + "condition_embedding.weight"]), pk.vec(pe_m)
    w.lnf_g, w.lnf_b = pk.vec(sd[p + "decoder.norm.weight"]), pk.vec(sd[p + "decoder.norm.bias"])
    w.w_out, w.b_out = mat2(sd[p + "latent_proj.weight"], w16, "w_out"), pk.vec(sd[p + "latent_proj.bias"])
    for x in range(len(STREAMS)):
        zcat, ycat = torch.cat(zx[x], 0), torch.cat(yx[x], 0)
        w.w_zx[x], w.a_zx[x], w.w_yx[x] = pk.mat(zcat), pk.vec(torch.cat(az[x], 0)), pk.mat(ycat)
        if pk16 is not None:
            w16.w_zx[x], w16.w_yx[x] = pk16.mat(zcat), pk16.mat(ycat)
    w.layers = C.cast(layers, C.POINTER(_lib.DenoiserLayer))
    out = {"struct": w, "layers": layers, "keep": pk.keep}
    if pk16 is not None:
        w16.d_model, w16.latent_dim, w16.n_tokens, w16.n_layers, w16.n_heads, w16.ff_size = d, lat, n_tokens, n_layers, n_heads, ff
        w16.precision, w16.pe_len = precision, pe_m.shape[0]
        w16.layers = C.cast(layers16, C.POINTER(_lib.DenoiserLayer))
        out.update({"struct16": w16, "layers16": layers16, "keep16": pk16.keep})
    return out


def pack_vae(sd: Dict[str, Tensor], prefix: str, n_layers: int, n_heads: int, ff: int, precision: int, device) -> dict:
    p = prefix
    # the handle created right after this call reads the same switch (cfb_vae_create)
    pk = _Packer(device, precision, f16=bool(_lib.lib().cfb_get_vae_f16()))
    d = sd[p + "body_final_layer.weight"].shape[1]
    nb = (n_layers - 1) // 2
    w = _lib.VaeWeights()
    w.d_model, w.n_layers, w.n_heads, w.ff_size, w.precision = d, n_layers, n_heads, ff, precision
    pe_q = sd[p + "query_pos_decoder.pe"].detach().float().cpu()[:, 0]
    pe_m = sd[p + "mem_pos_decoder.pe"].detach().float().cpu()[:, 0]
    w.pe_len = min(pe_q.shape[0], pe_m.shape[0])
    w.pe_query, w.pe_mem = pk.vec(pe_q), pk.vec(pe_m)
    arrays = []
    for pi, part in enumerate(("body", "hands")):
        dp = f"{p}{part}_decoder."
        names = [f"input_blocks.{i}." for i in range(nb)] + ["middle_block."] + [f"output_blocks.{i}." for i in range(nb)]
        layers = (_lib.VaeLayer * n_layers)()
        for li, nm in enumerate(names):
            lp, L = dp + nm, layers[li]
            L.ln1_g, L.ln1_b = pk.vec(sd[lp + "norm1.weight"]), pk.vec(sd[lp + "norm1.bias"])
            L.w_in, L.b_in = pk.mat(sd[lp + "self_attn.in_proj_weight"]), pk.vec(sd[lp + "self_attn.in_proj_bias"])
            L.w_so, L.b_so = pk.mat(sd[lp + "self_attn.out_proj.weight"]), pk.vec(sd[lp + "self_attn.out_proj.bias"])
            L.ln2_g, L.ln2_b = pk.vec(sd[lp + "norm2.weight"]), pk.vec(sd[lp + "norm2.bias"])
            W, b = sd[lp + "multihead_attn.in_proj_weight"], sd[lp + "multihead_attn.in_proj_bias"]
            L.w_q, L.b_q = pk.mat(W[:d]), pk.vec(b[:d])
            L.w_kv, L.b_kv = pk.mat(W[d:]), pk.vec(b[d:])
            L.w_co = pk.mat(sd[lp + "multihead_attn.out_proj.weight"])
            L.b_co = pk.vec(sd[lp + "multihead_attn.out_proj.bias"])
            L.ln3_g, L.ln3_b = pk.vec(sd[lp + "norm3.weight"]), pk.vec(sd[lp + "norm3.bias"])
            L.w_ff1, L.b_ff1 = pk.mat(sd[lp + "linear1.weight"]), pk.vec(sd[lp + "linear1.bias"])
            L.w_ff2, L.b_ff2 = pk.mat(sd[lp + "linear2.weight"]), pk.vec(sd[lp + "linear2.bias"])
        arrays.append(layers)
        D = w.part[pi]
        D.layers = C.cast(layers, C.POINTER(_lib.VaeLayer))
        for i in range(nb):
            D.w_skip[i] = pk.mat(sd[dp + f"linear_blocks.{i}.weight"])
            D.b_skip[i] = pk.vec(sd[dp + f"linear_blocks.{i}.bias"])
        D.lnf_g, D.lnf_b = pk.vec(sd[dp + "norm.weight"]), pk.vec(sd[dp + "norm.bias"])
        D.w_final = pk.mat(sd[p + f"{part}_final_layer.weight"])
        D.b_final = pk.vec(sd[p + f"{part}_final_layer.bias"])
        D.n_out = sd[p + f"{part}_final_layer.weight"].shape[0]
    # encode side (vae.py:162-266)
    w.pe_enc = pk.vec(sd[p + "query_pos_encoder.pe"].detach().float().cpu()[:, 0])
    col0 = 0
    for pi, part in enumerate(("body", "hands")):
        ep = f"{p}{part}_encoder."
        names = [f"input_blocks.{i}." for i in range(nb)] + ["middle_block."] + [f"output_blocks.{i}." for i in range(nb)]
        layers = (_lib.VaeEncLayer * n_layers)()
        for li, nm in enumerate(names):
            lp, L = ep + nm, layers[li]
            L.ln1_g, L.ln1_b = pk.vec(sd[lp + "norm1.weight"]), pk.vec(sd[lp + "norm1.bias"])
            L.w_in, L.b_in = pk.mat(sd[lp + "self_attn.in_proj_weight"]), pk.vec(sd[lp + "self_attn.in_proj_bias"])
            L.w_so, L.b_so = pk.mat(sd[lp + "self_attn.out_proj.weight"]), pk.vec(sd[lp + "self_attn.out_proj.bias"])
            L.ln2_g, L.ln2_b = pk.vec(sd[lp + "norm2.weight"]), pk.vec(sd[lp + "norm2.bias"])
            L.w_ff1, L.b_ff1 = pk.mat(sd[lp + "linear1.weight"]), pk.vec(sd[lp + "linear1.bias"])
            L.w_ff2, L.b_ff2 = pk.mat(sd[lp + "linear2.weight"]), pk.vec(sd[lp + "linear2.bias"])
        arrays.append(layers)
        E = w.enc[pi]
        E.layers = C.cast(layers, C.POINTER(_lib.VaeEncLayer))
        for i in range(nb):
            E.w_skip[i] = pk.mat(sd[ep + f"linear_blocks.{i}.weight"])
            E.b_skip[i] = pk.vec(sd[ep + f"linear_blocks.{i}.bias"])
        E.lnf_g, E.lnf_b = pk.vec(sd[ep + "norm.weight"]), pk.vec(sd[ep + "norm.bias"])
        E.tokens = pk.vec(sd[p + f"{part}_global_motion_token"])
        E.w_emb, E.b_emb = pk.vec(sd[p + f"{part}_skel_embedding.weight"]), pk.vec(sd[p + f"{part}_skel_embedding.bias"])
        E.n_in, E.col0 = sd[p + f"{part}_skel_embedding.weight"].shape[1], col0
        col0 += E.n_in
    return {"struct": w, "layers": arrays, "keep": pk.keep}
