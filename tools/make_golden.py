#!/usr/bin/env python
"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules from /root/reference.

Runs only in the build container (the GPU box has no /root/reference); the outputs are committed.
Weights and inputs come from convofusion_b200.synthetic (seeded torch-CPU RNG), so tests regenerate the
inputs from the same seeds and only outputs are stored.

What executes reference code:   Denoiser, ConvoFusionVae, AudioConvEncoder, TextAudioMotionFuser.
What is restated (cannot import: lightning/torchmetrics/kornia/nltk/diffusers missing):
  the 7-branch batch assembly + reverse loops (oracle/convofusion_oracle.py, line-by-line from
  convofusion.py:391-549,909-973 and unbounded_synthesis.py:28-187) and the DDIM/DDPM schedulers.
  In the "sample_*" goldens the reference Denoiser/VAE run INSIDE those restated loops.
  tools/pin_reference_loops.py closes that gap: it imports the reference's loop / test_diffusion_forward code itself
  (stand-in modules for the absent packages) and shows the restated loops are bit-identical to it (ref_loops.pt).

Usage: python tools/make_golden.py [--ref /root/reference]
"""
import argparse
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--ref", default="/root/reference")
args = ap.parse_args()
sys.path.insert(0, args.ref)
if "omegaconf" not in sys.modules:     # only convofusion/config.py needs it (instantiate_from_config)
    stub = types.ModuleType("omegaconf")
    stub.OmegaConf = object
    sys.modules["omegaconf"] = stub

from convofusion.models.architectures.denoiser import Denoiser as RefDenoiser          # noqa: E402
from convofusion.models.architectures.vae import ConvoFusionVae as RefVae              # noqa: E402
from convofusion.models.architectures.audioenc import AudioConvEncoder as RefAudio     # noqa: E402
from convofusion.models.architectures.condfuser import TextAudioMotionFuser as RefFuser  # noqa: E402

import convofusion_b200 as cf                                                          # noqa: E402
from convofusion_b200.synthetic import randomize_, synthetic_clip                      # noqa: E402
from oracle import convofusion_oracle as O                                             # noqa: E402

torch.manual_seed(1234)       # the reference's seed (configs/base.yaml:2)
torch.set_num_threads(8)
OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)

abl = types.SimpleNamespace(SKIP_CONNECT=True, VAE_TYPE="convofusion", DIFF_PE_TYPE="convofusion", CAUSAL_ATTN=False,
                            MLP_DIST=False, PE_TYPE="convofusion")
sampler = randomize_(cf.ConvoFusionSampler(precision="fp32"), 1234)
sd = {k: v.clone() for k, v in sampler.state_dict().items()}

ref_den = RefDenoiser(ablation=abl, nfeats=189, condition="text+audio", latent_dim=[1, 128], ff_size=1024, num_layers=9,
                      num_heads=4, dropout=0.1, normalize_before=True, activation="gelu", flip_sin_to_cos=True,
                      return_intermediate_dec=False, position_embedding="sine", arch="trans_dec", freq_shift=0,
                      text_encoded_dim=512, audio_encoded_dim=512).eval()
ref_vae = RefVae(ablation=abl, nfeats=189, latent_dim=[1, 128], ff_size=1024, num_layers=5, num_heads=2, dropout=0.1,
                 arch="encoder_decoder", normalize_before=True, activation="gelu", position_embedding="sine").eval()
ref_audio = RefAudio(80, 256, 512, max_seq_len=128, fps=25, sample_rate=16000, hop_length=512).eval()
ref_fuser = RefFuser(types.SimpleNamespace(model=types.SimpleNamespace(latent_dim=[1, 128], vae_type="convofusion")), 512).eval()
ref_textproj = torch.nn.Sequential(torch.nn.ReLU(), torch.nn.Linear(768, 512)).eval()    # t5.py:48-49
print("strict loads:",
      ref_den.load_state_dict(sampler.denoiser.state_dict(), strict=True),
      ref_vae.load_state_dict(sampler.vae.state_dict(), strict=True),
      ref_audio.load_state_dict(sampler.text_audio_encoder.audio_encoder.state_dict(), strict=True),
      ref_fuser.load_state_dict(sampler.condition_fuser.state_dict(), strict=True),
      ref_textproj.load_state_dict({"1.weight": sd["text_audio_encoder.text_encoder.projection.1.weight"],
                                    "1.bias": sd["text_audio_encoder.text_encoder.projection.1.bias"]}))
assert ref_audio.audio_max_length == 161


def ref_guidance_batch(syn):
    """convofusion.py:909-973 with the reference's own modules on featurised text."""
    clip, U, Ua = syn["clip"], syn["uncond_text"], syn["uncond_text_attn"]
    B = clip["mel_lsn"].shape[0]
    Ub, Uab = U[None].expand(B, -1, -1), Ua[None].expand(B, -1)
    tl, tla, ts, tsa = clip["text_lsn"], clip["text_lsn_attn"], clip["text_spk"], clip["text_spk_attn"]
    text_lsn = torch.cat([Ub, tl, Ub, Ub, Ub, Ub, tl]); attn_lsn = torch.cat([Uab, tla, Uab, Uab, Uab, Uab, tla])
    text_spk = torch.cat([Ub, Ub, Ub, ts, Ub, Ub, ts]); attn_spk = torch.cat([Uab, Uab, Uab, tsa, Uab, Uab, tsa])
    mel = clip["mel_lsn"]
    um = -90 * torch.ones_like(mel); um[..., 40:45] = 0
    mel7 = torch.cat([um, um, mel, um, um, um, mel])
    apb = clip["apb"]; two = 2 * torch.ones_like(apb)
    apb7 = torch.cat([two, two, two, two, apb, two, apb])
    ids = list(clip["lsn_id"]); ids7 = [0] * (5 * B) + ids + ids
    with torch.no_grad():
        tspk, tlsn, alsn = ref_textproj(text_spk), ref_textproj(text_lsn), ref_audio(mel7)
        enc = ref_fuser(tspk, alsn, tlsn, apb7, ids7)
    masks = {"alsn": None, "tlsn": ~attn_lsn.bool(), "spkemb": ~attn_spk.bool()}    # audioenc.py:61
    return enc, masks


def ref_denoise(x, t, enc, masks):
    with torch.no_grad():
        return ref_den(sample=x, timestep=torch.as_tensor(t), encoder_hidden_states=list(enc), lengths=None,
                       mem_mask_dict=masks)


def save(name, obj):
    torch.save(obj, OUT / name)
    print(f"{name}: {(OUT / name).stat().st_size / 1024:.0f} KiB")


# 1 -- conditioning projections ------------------------------------------------------------------
g = torch.Generator().manual_seed(7)
mel_s = torch.rand(2, 24, 80, generator=g) * 80 - 80
t5_s = torch.randn(2, 6, 768, generator=g)
with torch.no_grad():
    save("conditioning.pt", {"audio": ref_audio(mel_s), "text": ref_textproj(t5_s),
                             "fuser_apb": ref_fuser(t5_s, t5_s, t5_s, torch.tensor([[0, 1, 2, 1]]), [3])[3],
                             "fuser_id": ref_fuser(t5_s, t5_s, t5_s, torch.tensor([[0, 1, 2, 1]]), [3, 35, 0])[4]})

# 2 -- one denoiser evaluation: monadic B=1 and dyadic B=2 ---------------------------------------
for tag, B, dyadic, t in (("mono_b1", 1, False, 481), ("dyad_b2", 2, True, 37)):
    syn = synthetic_clip(B, seed=1234 + B, dyadic=dyadic)
    enc, masks = ref_guidance_batch(syn)
    x = torch.randn(B, 16, 128, generator=torch.Generator().manual_seed(99 + B))
    eps, att = ref_denoise(torch.cat([x] * 7), t, enc, masks)
    save(f"denoiser_{tag}.pt", {"t": t, "eps": eps, "att_full": [a.chunk(7)[-1].clone() for a in att],
                                "enc_checksum": [float(e.double().sum()) for e in enc]})

# 3 -- VAE decode, ragged lengths ----------------------------------------------------------------
z = torch.randn(2, 3, 8, 128, generator=torch.Generator().manual_seed(5))
with torch.no_grad():
    save("vae_decode.pt", {"lengths": [128, 100, 37], "out": ref_vae.decode(z, [128, 100, 37])})
    save("vae_decode_short.pt", {"lengths": [64, 33], "out": ref_vae.decode(z[:, :2], [64, 33])})

# 3b - VAE encode (distribution parameters + root-subtracted features), ragged lengths ------------
xf = torch.randn(3, 128, 189, generator=torch.Generator().manual_seed(6))
with torch.no_grad():
    _, dist_e, feats_e = ref_vae.encode(xf, [128, 100, 37])
    _, dist_s, feats_s = ref_vae.encode(xf[:2, :32], [32, 5])
    save("vae_encode.pt", {"lengths": [128, 100, 37], "mu": dist_e.loc, "std": dist_e.scale, "feats": feats_e,
                           "short_lengths": [32, 5], "short_mu": dist_s.loc, "short_std": dist_s.scale})

# 4 -- full sampling run, config 1 of BASELINE.json: B=1, DDIM-50, guidance 7.5, then decode ------
syn = synthetic_clip(1, seed=1235, dyadic=False)
enc, masks = ref_guidance_batch(syn)
init = torch.randn(1, 16, 128, generator=torch.Generator().manual_seed(100))
for tag, kw in (("clip", dict(clip_sample=True)), ("mld", dict(clip_sample=False, set_alpha_to_one=False, steps_offset=1))):
    sch = O.DDIMSchedulerOracle(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                                beta_schedule="scaled_linear", **kw)
    rec = []
    zfin, att = O.diffusion_reverse(ref_denoise, sch, enc, masks, init, 50, guidance_scale=7.5, eta=0.0, record=rec)
    with torch.no_grad():
        joints = ref_vae.decode(O.latents_to_vae_input(zfin), [128])
    save(f"sample_ddim50_{tag}.pt", {"record": torch.stack(rec), "joints": joints,
                                     "att_last_tlsn": att[int(sch.timesteps[-1])][2]})

# 5 -- DDPM (the shipped scheduler), 10 steps, fixed step noise -----------------------------------
sch = O.DDPMSchedulerOracle(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                            clip_sample=True)
noise = torch.randn(10, 1, 16, 128, generator=torch.Generator().manual_seed(101))
rec = []
O.diffusion_reverse(ref_denoise, sch, enc, masks, init, 10, guidance_scale=7.5, step_noise=noise, record=rec)
save("sample_ddpm10.pt", {"record": torch.stack(rec)})

# 6 -- unbounded synthesis: 3 windows, B=2 dyadic, 6 DDIM steps, preseq inpainting + root stitching -
sch = O.DDIMSchedulerOracle(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                            clip_sample=True)
nsch = O.DDPMSchedulerOracle(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                             clip_sample=True)
preseq, prev, feats_all, z_all = None, None, [], []
for k in range(3):
    syn = synthetic_clip(2, seed=2000 + k, dyadic=True)
    enc, masks = ref_guidance_batch(syn)
    init = torch.randn(2, 16, 128, generator=torch.Generator().manual_seed(300 + k))
    zk, _ = O.diffusion_reverse_forecast(ref_denoise, sch, nsch, enc, masks, init, 6, preseq, guidance_scale=7.5)
    preseq = zk[zk.shape[0] // 2:].permute(1, 0, 2).clone()                    # unbounded_synthesis.py:442-444
    with torch.no_grad():
        feats = ref_vae.decode(O.latents_to_vae_input(zk), [128, 128])
    feats = O.stitch_root(feats, prev)
    prev = feats[:, 64:, :]
    feats_all.append(feats); z_all.append(zk)
save("unbounded_3win.pt", {"z": torch.stack(z_all), "feats": torch.stack(feats_all)})
print("done")
