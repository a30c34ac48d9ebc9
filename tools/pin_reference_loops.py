#!/usr/bin/env python
"""Pin the oracle's two reverse loops against the reference's OWN loop code.

`Convofusion._diffusion_reverse` (convofusion/models/modeltype/convofusion.py:391-549) and
`diffusion_reverse_forecast` (unbounded_synthesis.py:28-187) live in modules whose imports (pytorch_lightning,
torchmetrics, kornia, nltk, omegaconf, soundfile, matplotlib, diffusers ...) are not installed here.  Only NAMES
from those packages are needed to import the two modules, so this script registers empty stand-in modules for
them, imports the unmodified reference sources from /root/reference and calls the two functions with a
stand-in `self` / `model` that carries
    denoiser         the unmodified reference Denoiser with the golden weights
    scheduler        oracle DDIM (diffusers 0.14.0 itself is absent: that arithmetic stays "parity unpinned")
    noise_scheduler  oracle DDPM (add_noise)
It then (1) asserts that the oracle's restated loops, driven by the SAME reference denoiser and the same global-RNG
noise, reproduce the reference loops bit for bit -- loop structure, 7-way guidance combine, attention-map selection,
latent inpainting incl. the aliasing quirk -- and (2) stores the reference loops' outputs in tests/golden/ref_loops.pt
for tests/test_oracle.py, which replays them with the oracle's own denoiser (no /root/reference at test time).

Usage (build container only): python tools/pin_reference_loops.py [--ref /root/reference]
"""
import argparse
import importlib
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


class _Names(types.ModuleType):
    """Stand-in for an absent package: every attribute is an inert class, enough for `from x import Y`."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


def stand_in(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _Names(n)
            m.__path__ = []
            sys.modules[n] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)


def import_reference(module):
    for _ in range(40):
        try:
            return importlib.import_module(module)
        except ModuleNotFoundError as exc:
            if exc.name.startswith("convofusion."):
                raise
            print(f"  stand-in for absent package: {exc.name}")
            stand_in(exc.name)
    raise RuntimeError(f"could not import {module}")


import numpy as np                                                              # noqa: E402
for _alias, _ty in (("float", float), ("int", int), ("bool", bool), ("object", object)):
    if not hasattr(np, _alias):      # the reference's data utilities (imported by the script, unused here) predate numpy 1.24
        setattr(np, _alias, _ty)
stand_in("pytorch_lightning")
sys.modules["pytorch_lightning"].LightningModule = torch.nn.Module
ref_model = import_reference("convofusion.models.modeltype.convofusion")
# the script imports `convofusion.models.tools.weg`, a module name that does not exist in the reference tree
# (the file is word_excitation_guidance.py): same shim as INTEGRATION.md section 4
import convofusion.models.tools as _tools                                       # noqa: E402
_weg = importlib.import_module("convofusion.models.tools.word_excitation_guidance")
sys.modules["convofusion.models.tools.weg"] = _weg
_tools.weg = _weg
ref_script = import_reference("unbounded_synthesis")
from convofusion.models.architectures.denoiser import Denoiser as RefDenoiser   # noqa: E402

import convofusion_b200 as cf                                                   # noqa: E402
from convofusion_b200.synthetic import randomize_, synthetic_clip               # noqa: E402
from oracle import convofusion_oracle as O                                      # noqa: E402

torch.set_num_threads(8)
abl = types.SimpleNamespace(SKIP_CONNECT=True, VAE_TYPE="convofusion", DIFF_PE_TYPE="convofusion", CAUSAL_ATTN=False,
                            MLP_DIST=False, PE_TYPE="convofusion")
sampler = randomize_(cf.ConvoFusionSampler(precision="fp32"), 1234)
sd = {k: v.clone() for k, v in sampler.state_dict().items()}
ref_den = RefDenoiser(ablation=abl, nfeats=189, condition="text+audio", latent_dim=[1, 128], ff_size=1024, num_layers=9,
                      num_heads=4, dropout=0.1, normalize_before=True, activation="gelu", flip_sin_to_cos=True,
                      return_intermediate_dec=False, position_embedding="sine", arch="trans_dec", freq_shift=0,
                      text_encoded_dim=512, audio_encoded_dim=512).eval()
print("strict load:", ref_den.load_state_dict(sampler.denoiser.state_dict(), strict=True))

SCHED = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
N_STEPS, B, SEED = 6, 2, 4242


def stand_in_model(scheduler):
    cfg = types.SimpleNamespace(model=types.SimpleNamespace(
        scheduler=types.SimpleNamespace(num_inference_timesteps=N_STEPS, eta=0.0)), DATASET=types.SimpleNamespace(NFEATS=189))
    return types.SimpleNamespace(
        weg_parameters={"scale_range": (1.0, 0.5), "thresholds": {}}, do_classifier_free_guidance=True,
        clf_guidance_drops=6, vae_type="convofusion", latent_dim=[1, 128], guidance_scale=7.5, cfg=cfg,
        scheduler=scheduler, noise_scheduler=O.DDPMSchedulerOracle(clip_sample=True, **SCHED), denoiser=ref_den)


def ref_denoise(x, t, enc, masks):
    return ref_den(sample=x, timestep=torch.as_tensor(t), encoder_hidden_states=list(enc), lengths=None, mem_mask_dict=masks)


def seven_branch_batch(seed):
    syn = synthetic_clip(B, seed=seed, dyadic=True)
    clip = dict(syn["clip"])
    clip["text_lsn_mask"], clip["text_spk_mask"] = ~clip["text_lsn_attn"].bool(), ~clip["text_spk_attn"].bool()
    return O.assemble_guidance_batch(sd, clip, syn["uncond_text"], ~syn["uncond_text_attn"].bool())


out = {"n_steps": N_STEPS, "B": B, "seed": SEED}
with torch.no_grad():
    # ---- Convofusion._diffusion_reverse, both DDIM parameterisations of SURVEY 8d
    for tag, kw in (("clip", dict(clip_sample=True)), ("mld", dict(clip_sample=False, set_alpha_to_one=False, steps_offset=1))):
        enc, masks = seven_branch_batch(3100)
        model = stand_in_model(O.DDIMSchedulerOracle(**SCHED, **kw))
        torch.manual_seed(SEED)
        z_ref, att_ref = ref_model.Convofusion._diffusion_reverse(model, list(enc), lengths=[128] * B, cond_masks=masks)
        torch.manual_seed(SEED)
        init = torch.randn(B, 16, 128)
        rec = []
        z_or, att_or = O.diffusion_reverse(ref_denoise, O.DDIMSchedulerOracle(**SCHED, **kw), enc, masks, init, N_STEPS,
                                           guidance_scale=7.5, eta=0.0, record=rec)
        assert torch.equal(z_ref, z_or), f"_diffusion_reverse[{tag}]: oracle loop differs from the reference loop"
        assert sorted(att_ref) == sorted(att_or)
        for t in att_ref:
            for a, b in zip(att_ref[t], att_or[t]):
                assert torch.equal(a, b), "attention maps of the full-cond branch differ"
        t_last = sorted(att_ref)[0]
        out[f"reverse_{tag}"] = {"z": z_ref.clone(), "att_last_tlsn": att_ref[t_last][2].clone(), "t_last": int(t_last)}
        print(f"_diffusion_reverse[{tag}]: oracle loop == reference loop (bit for bit), |z|max {float(z_ref.abs().max()):.3f}")

    # ---- diffusion_reverse_forecast: first window (no preseq) and a window that inpaints 8 latent tokens
    enc, masks = seven_branch_batch(3200)
    preseq = torch.randn(B, 8, 128, generator=torch.Generator().manual_seed(77))
    for tag, pre in (("first", None), ("inpaint", preseq)):
        model = stand_in_model(O.DDIMSchedulerOracle(clip_sample=True, **SCHED))
        torch.manual_seed(SEED + 1)
        z_ref, att_ref = ref_script.diffusion_reverse_forecast(model, list(enc), lengths=[128] * B, preseq=pre, cond_masks=masks)
        torch.manual_seed(SEED + 1)
        init = torch.randn(B, 16, 128)
        z_or, att_or = O.diffusion_reverse_forecast(ref_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED),
                                                    O.DDPMSchedulerOracle(clip_sample=True, **SCHED), enc, masks, init,
                                                    N_STEPS, pre, guidance_scale=7.5)
        assert torch.equal(z_ref, z_or), f"diffusion_reverse_forecast[{tag}]: oracle loop differs from the reference loop"
        for a, b in zip(att_ref, att_or):
            assert torch.equal(a, b)
        out[f"forecast_{tag}"] = {"z": z_ref.clone(), "att_last_alsn": att_ref[1].clone()}
        print(f"diffusion_reverse_forecast[{tag}]: oracle loop == reference loop (bit for bit)")
    out["preseq_seed"] = 77

# ---- Convofusion.test_diffusion_forward (convofusion.py:817-1065): the 7-branch batch assembly, both text/audio
# encoders' forward code (TextAudioController.forward audioenc.py:52-91, T5TextEncoder.forward t5.py:51-59 over a
# stand-in for the frozen T5 body: text string -> synthetic last hidden state), the condition fuser, the reverse
# loop, the latent reshape and ConvoFusionVae.decode -- all reference code, on a stand-in `self`.
from convofusion.models.architectures.vae import ConvoFusionVae as RefVae              # noqa: E402
from convofusion.models.architectures.audioenc import AudioConvEncoder as RefAudio     # noqa: E402
from convofusion.models.architectures.audioenc import TextAudioController as RefController  # noqa: E402
from convofusion.models.architectures.condfuser import TextAudioMotionFuser as RefFuser  # noqa: E402
from convofusion.models.architectures.t5 import T5TextEncoder as RefT5                 # noqa: E402

ref_vae = RefVae(ablation=abl, nfeats=189, latent_dim=[1, 128], ff_size=1024, num_layers=5, num_heads=2, dropout=0.1,
                 arch="encoder_decoder", normalize_before=True, activation="gelu", position_embedding="sine").eval()
ref_audio = RefAudio(80, 256, 512, max_seq_len=128, fps=25, sample_rate=16000, hop_length=512).eval()
ref_fuser = RefFuser(types.SimpleNamespace(model=types.SimpleNamespace(latent_dim=[1, 128], vae_type="convofusion")), 512).eval()
ref_proj = torch.nn.Sequential(torch.nn.ReLU(), torch.nn.Linear(768, 512)).eval()            # t5.py:48-49
print("strict loads:", ref_vae.load_state_dict(sampler.vae.state_dict(), strict=True),
      ref_audio.load_state_dict(sampler.text_audio_encoder.audio_encoder.state_dict(), strict=True),
      ref_fuser.load_state_dict(sampler.condition_fuser.state_dict(), strict=True),
      ref_proj.load_state_dict({"1.weight": sd["text_audio_encoder.text_encoder.projection.1.weight"],
                                "1.bias": sd["text_audio_encoder.text_encoder.projection.1.bias"]}))

syn = synthetic_clip(B, seed=3300, dyadic=True)
clip = syn["clip"]
t5_table = {"-" * 10: (syn["uncond_text"], syn["uncond_text_attn"])}
for i in range(B):
    t5_table[f"listener utterance {i}"] = (clip["text_lsn"][i], clip["text_lsn_attn"][i])
    t5_table[f"speaker utterance {i}"] = (clip["text_spk"][i], clip["text_spk_attn"][i])


def t5_body_stand_in(texts, return_map=False):
    """get_last_hidden_state (t5.py:87-108) without t5-base: (last hidden state, attention mask as bool, word map)."""
    hid = torch.stack([t5_table[t][0] for t in texts])
    attn = torch.stack([t5_table[t][1] for t in texts]).to(dtype=bool)
    return hid, attn, ([t.split() for t in texts] if return_map else None)


t5_self = types.SimpleNamespace(get_last_hidden_state=t5_body_stand_in, projection=ref_proj)
controller_self = types.SimpleNamespace(text_encoder=lambda texts, return_map=False: RefT5.forward(t5_self, texts, return_map),
                                        audio_encoder=ref_audio)
model = stand_in_model(O.DDIMSchedulerOracle(clip_sample=True, **SCHED))
model.condition, model.WEG_type, model.vae, model.condition_fuser = "text+audio", "no", ref_vae, ref_fuser
model.text_audio_encoder = lambda text, audio, person_type, return_textmap=False: RefController.forward(
    controller_self, text, audio, person_type, return_textmap)
model._diffusion_reverse = lambda *a, **k: ref_model.Convofusion._diffusion_reverse(model, *a, **k)
motion = torch.randn(B, 128, 189, generator=torch.Generator().manual_seed(78))
lengths = [128, 100]
batch = {"length": lengths, "text_lsn": [f"listener utterance {i}" for i in range(B)],
         "text_spk": [f"speaker utterance {i}" for i in range(B)], "melspec_lsn": clip["mel_lsn"],
         "melspec_spk": clip["mel_lsn"].flip(0), "active_passive_lsn": clip["apb"], "motion_spk": motion,
         "lsn_id": list(clip["lsn_id"]), "motion_lsn": motion}
with torch.no_grad():
    torch.manual_seed(SEED + 2)
    rs = ref_model.Convofusion.test_diffusion_forward(model, batch)
    # the oracle's restatement of the same call chain, with the reference Denoiser / VAE inside
    oc = dict(clip)
    oc["text_lsn_mask"], oc["text_spk_mask"] = ~clip["text_lsn_attn"].bool(), ~clip["text_spk_attn"].bool()
    enc, masks = O.assemble_guidance_batch(sd, oc, syn["uncond_text"], ~syn["uncond_text_attn"].bool())
    torch.manual_seed(SEED + 2)
    init = torch.randn(B, 16, 128)
    z_or, _ = O.diffusion_reverse(ref_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED), enc, masks, init, N_STEPS,
                                  guidance_scale=7.5)
    joints_or = ref_vae.decode(O.latents_to_vae_input(z_or), lengths)
lat_or = O.latents_to_vae_input(z_or).permute(1, 2, 0, 3)
d_lat = float((rs["lat_t"] - lat_or).abs().max() / lat_or.abs().max())
d_j = float((rs["m_rst"] - joints_or).abs().max() / joints_or.abs().max())
print(f"test_diffusion_forward: oracle chain vs reference call chain: latents {d_lat:.2e}, joints {d_j:.2e} (max-rel)")
assert d_lat < 2e-5 and d_j < 2e-5, "oracle's conditioning assembly / call chain differs from test_diffusion_forward"
assert rs["m_rst"].shape == (B, 128, 189) and rs["lat_m"].shape == rs["lat_t"].shape
out["forward"] = {"m_rst": rs["m_rst"].clone(), "lat_t": rs["lat_t"].clone(), "lengths": lengths, "clip_seed": 3300}

# ---- process_samples (unbounded_synthesis.py:244-512): serial windows at 50 % overlap over a long batch -- window
# slicing of mel / active-passive bits, text by timestamp (process_text :189-241), 7-branch assembly per window,
# diffusion_reverse_forecast with the previous window's last 8 latent tokens, decode, root x/z stitching.
from convofusion_b200.synthetic import synthetic_long_batch, synthetic_text_features   # noqa: E402

N_PARTS = 2
long_batch = synthetic_long_batch(B, N_PARTS, seed=3400)
uncond = (syn["uncond_text"], syn["uncond_text_attn"])
encoder_calls = []


def t5_body_any_string(texts, return_map=False):
    encoder_calls.append(list(texts))
    feats = [uncond if t == "-" * 10 else synthetic_text_features(t) for t in texts]
    hid, attn = torch.stack([f[0] for f in feats]), torch.stack([f[1] for f in feats]).to(dtype=bool)
    return hid, attn, ([t.split() for t in texts] if return_map else None)


t5_self2 = types.SimpleNamespace(get_last_hidden_state=t5_body_any_string, projection=ref_proj)
controller_self2 = types.SimpleNamespace(text_encoder=lambda texts, return_map=False: RefT5.forward(t5_self2, texts, return_map),
                                         audio_encoder=ref_audio)
saved = []
model = stand_in_model(O.DDIMSchedulerOracle(clip_sample=True, **SCHED))
model.condition, model.WEG_type, model.vae, model.condition_fuser = "text+audio", "no", ref_vae, ref_fuser
model.device = torch.device("cpu")
model.text_audio_encoder = lambda text, audio, person_type, return_textmap=False: RefController.forward(
    controller_self2, text, audio, person_type, return_textmap)
model.save_npy = lambda tup: saved.append(tup[1].clone())          # feats_rst of the window, after stitching
with torch.no_grad():
    torch.manual_seed(SEED + 3)
    ref_script.process_samples(long_batch, model, None, None, None)
n_windows = 2 * N_PARTS - 1
assert len(saved) == n_windows and len(encoder_calls) == 2 * n_windows
texts_spk = [encoder_calls[2 * k][3 * B:4 * B] for k in range(n_windows)]      # spk-only branch rows
texts_lsn = [encoder_calls[2 * k + 1][B:2 * B] for k in range(n_windows)]      # text-only branch rows
print("window texts (stream 1):", [t[1][:40] for t in texts_lsn])

# the oracle's restatement of the same driver (what ConvoFusionSampler.synthesize_unbounded mirrors), reference modules inside
with torch.no_grad():
    torch.manual_seed(SEED + 3)
    preseq, prev, feats_or = None, None, []
    for k in range(n_windows):
        mel = long_batch["melspec_lsn"][:, int(k / 2 * 160):int((k / 2 + 1) * 160) + 1]
        apb = long_batch["active_passive_lsn"][:, int(k / 2 * 8):int((k / 2 + 1) * 8)]
        fl = [uncond if t == "-" * 10 else synthetic_text_features(t) for t in texts_lsn[k]]
        fs = [uncond if t == "-" * 10 else synthetic_text_features(t) for t in texts_spk[k]]
        wclip = {"mel_lsn": mel, "apb": apb, "lsn_id": list(long_batch["lsn_id"]),
                 "text_lsn": torch.stack([f[0] for f in fl]), "text_lsn_attn": torch.stack([f[1] for f in fl]),
                 "text_spk": torch.stack([f[0] for f in fs]), "text_spk_attn": torch.stack([f[1] for f in fs])}
        wclip["text_lsn_mask"], wclip["text_spk_mask"] = ~wclip["text_lsn_attn"].bool(), ~wclip["text_spk_attn"].bool()
        enc, masks = O.assemble_guidance_batch(sd, wclip, syn["uncond_text"], ~syn["uncond_text_attn"].bool())
        init = torch.randn(B, 16, 128)
        zk, _ = O.diffusion_reverse_forecast(ref_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED),
                                             O.DDPMSchedulerOracle(clip_sample=True, **SCHED), enc, masks, init, N_STEPS,
                                             preseq, guidance_scale=7.5)
        preseq = zk[zk.shape[0] // 2:].permute(1, 0, 2).clone()
        feats = O.stitch_root(ref_vae.decode(O.latents_to_vae_input(zk), [128] * B), prev)
        prev = feats[:, 64:, :]
        feats_or.append(feats)
for k in range(n_windows):
    dk = float((saved[k] - feats_or[k]).abs().max())
    print(f"process_samples window {k}: oracle driver vs reference driver max |diff| {dk:.2e}")
    assert dk == 0.0, "oracle's window driver differs from process_samples"
# ---- the package's host-side window bookkeeping (convofusion_b200/windows.py) against process_text / process_samples
from convofusion_b200.windows import slice_windows, window_spans, window_text           # noqa: E402
import random                                                                           # noqa: E402

wins = slice_windows(long_batch, lambda texts: t5_body_any_string(texts)[:2])
assert len(wins) == n_windows
for k, w in enumerate(wins):
    assert w["texts"]["lsn"] == texts_lsn[k] and w["texts"]["spk"] == texts_spk[k], "window texts differ from process_samples"
rng = random.Random(5)
text_cases = []
for case in range(60):                      # random word timings: gaps, overlaps, words longer than a window
    t, segs = rng.uniform(0, 1.5), []
    while t < 16.0:
        dur = rng.choice([0.2, 0.4, 0.9, 2.5, 6.5])
        segs.append(((round(t, 3), round(t + dur, 3)), f"w{len(segs)}"))
        t += dur * rng.uniform(0.3, 1.4)
    for (t0, t1) in window_spans(3 * 128):
        ref_txt = ref_script.process_text([segs, "-" * 10], t0, t1)
        assert [window_text(segs, t0, t1), window_text("-" * 10, t0, t1)] == ref_txt, "window_text differs from process_text"
        if case < 12:
            text_cases.append({"segments": segs, "t0": t0, "t1": t1, "text": ref_txt[0]})
print(f"window_text == process_text on 60 random transcripts x 5 windows; slice_windows texts == process_samples")
out["window_text_cases"] = text_cases
out["unbounded"] = {"feats": torch.stack(saved), "texts_lsn": texts_lsn, "texts_spk": texts_spk, "n_parts": N_PARTS,
                    "batch_seed": 3400, "uncond_clip_seed": 3300}

# ---- word-excitation guidance (convofusion.py:437-496 + iterative_refinement_step :298-388 + tools/
# word_excitation_guidance.py): the reference loop with focus tokens, B = 1 (its own assertion), autograd through the
# reference Denoiser.  Two settings: plain latent updates on every step, and a threshold at step 0 that triggers the
# iterative refinement.  The oracle's restated loop must reproduce the reference bit for bit with the same denoiser.
def one_clip_batch(seed):
    syn1 = synthetic_clip(1, seed=seed, dyadic=True)
    c1 = dict(syn1["clip"])
    c1["text_lsn_mask"], c1["text_spk_mask"] = ~c1["text_lsn_attn"].bool(), ~c1["text_spk_attn"].bool()
    return O.assemble_guidance_batch(sd, c1, syn1["uncond_text"], ~syn1["uncond_text_attn"].bool())


weg_out = {"n_steps": N_STEPS, "clip_seed": 3500, "init_seed": SEED + 4, "cases": {}}
enc1, masks1 = one_clip_batch(3500)
eot = int(torch.argmax(masks1["tlsn"].chunk(7)[1].int(), dim=1)[0]) - 1
focus = [[2, max(3, min(5, eot - 2))]]
print(f"WEG: text-only branch has its EOS at token {eot}; focus tokens {focus}")
# Random-init weights give nearly uniform text attention (loss = 1 - 1/18 at every step) and gradients of ~1e-6, so the
# step size is scaled up until the guidance moves the latents by O(1) -- the code path does not depend on it.
for tag, wp in (("update", {"scale_range": [1.0, 0.5], "thresholds": {}, "scale_factor": 2.0e7, "max_iter_to_alter": 4,
                            "max_refinement_steps": 3}),
                ("refine", {"scale_range": [1.0, 0.5], "thresholds": {0: 0.06, 2: 0.5}, "scale_factor": 1.0e7,
                            "max_iter_to_alter": 25, "max_refinement_steps": 3})):
    model = stand_in_model(O.DDIMSchedulerOracle(clip_sample=True, **SCHED))
    model.weg_parameters = {k: (dict(v) if isinstance(v, dict) else (list(v) if isinstance(v, list) else v)) for k, v in wp.items()}
    model.iterative_refinement_step = types.MethodType(ref_model.Convofusion.iterative_refinement_step, model)
    torch.manual_seed(SEED + 4)
    z_ref, att_ref = ref_model.Convofusion._diffusion_reverse(model, list(enc1), lengths=[128], cond_masks=masks1,
                                                               focus_indices=[list(f) for f in focus])
    torch.manual_seed(SEED + 4)
    init = torch.randn(1, 16, 128)
    log = []
    z_or, att_or = O.diffusion_reverse(ref_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED), enc1, masks1, init,
                                       N_STEPS, guidance_scale=7.5, focus_indices=focus,
                                       weg={k: (dict(v) if isinstance(v, dict) else v) for k, v in wp.items()}, weg_log=log)
    assert torch.equal(z_ref.detach(), z_or.detach()), f"WEG[{tag}]: oracle loop differs from the reference loop"
    # the same run without focus tokens must differ (the guidance really moved the latents)
    with torch.no_grad():
        z_plain, _ = O.diffusion_reverse(ref_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED), enc1, masks1, init,
                                         N_STEPS, guidance_scale=7.5)
    moved = float((z_or.detach() - z_plain).norm() / z_plain.norm())
    print(f"WEG[{tag}]: oracle loop == reference loop (bit for bit); losses {[round(e['loss'], 4) for e in log]}, "
          f"refinement iterations {[e['n_refine'] for e in log]}, moved the result by {moved:.3f} (L2)")
    assert moved > 1e-3
    weg_out["cases"][tag] = {"z": z_ref.detach().clone(), "params": wp, "focus": focus, "log": log}
# the same block inside diffusion_reverse_forecast (unbounded_synthesis.py:78-142: hard-coded parameters, scale factor
# 100, so with random-init weights the update is ~1e-4 of the latents -- the bit-for-bit comparison still sees it)
pre1 = torch.randn(1, 8, 128, generator=torch.Generator().manual_seed(79))
model = stand_in_model(O.DDIMSchedulerOracle(clip_sample=True, **SCHED))
model.iterative_refinement_step = types.MethodType(ref_model.Convofusion.iterative_refinement_step, model)
torch.manual_seed(SEED + 5)
z_ref, _ = ref_script.diffusion_reverse_forecast(model, list(enc1), lengths=[128], preseq=pre1, cond_masks=masks1,
                                                 focus_indices=[list(f) for f in focus])
torch.manual_seed(SEED + 5)
init = torch.randn(1, 16, 128)
log = []
z_or, _ = O.diffusion_reverse_forecast(ref_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED),
                                       O.DDPMSchedulerOracle(clip_sample=True, **SCHED), enc1, masks1, init, N_STEPS, pre1,
                                       guidance_scale=7.5, focus_indices=focus, weg_log=log)
assert torch.equal(z_ref.detach(), z_or.detach()), "WEG[forecast]: oracle loop differs from the reference loop"
with torch.no_grad():
    z_plain, _ = O.diffusion_reverse_forecast(ref_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED),
                                              O.DDPMSchedulerOracle(clip_sample=True, **SCHED), enc1, masks1, init, N_STEPS,
                                              pre1, guidance_scale=7.5)
moved = float((z_or.detach() - z_plain).norm() / z_plain.norm())
print(f"WEG[forecast]: oracle loop == reference diffusion_reverse_forecast (bit for bit); moved the result by {moved:.2e}")
assert 0 < moved
weg_out["cases"]["forecast"] = {"z": z_ref.detach().clone(), "params": dict(O.FORECAST_WEG), "focus": focus, "log": log,
                                "preseq_seed": 79, "init_seed": SEED + 5}
# ... and with a step size that makes the guidance matter (the oracle's loop takes the parameters as an argument)
big = dict(O.FORECAST_WEG, scale_factor=1.0e7, thresholds={0: 0.06}, max_refinement_steps=2)
log = []
z_big, _ = O.diffusion_reverse_forecast(ref_denoise, O.DDIMSchedulerOracle(clip_sample=True, **SCHED),
                                        O.DDPMSchedulerOracle(clip_sample=True, **SCHED), enc1, masks1, init, N_STEPS, pre1,
                                        guidance_scale=7.5, focus_indices=focus, weg=big, weg_log=log)
print(f"WEG[forecast_big]: moved the result by {float((z_big.detach() - z_plain).norm() / z_plain.norm()):.3f}; "
      f"refinement iterations {[e['n_refine'] for e in log]}")
weg_out["cases"]["forecast_big"] = {"z": z_big.detach().clone(), "params": big, "focus": focus, "log": log,
                                    "preseq_seed": 79, "init_seed": SEED + 5}
torch.save(weg_out, ROOT / "tests" / "golden" / "ref_weg.pt")

path = ROOT / "tests" / "golden" / "ref_loops.pt"
torch.save(out, path)
print(f"{path.name}: {path.stat().st_size / 1024:.0f} KiB")
