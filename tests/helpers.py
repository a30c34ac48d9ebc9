"""Shared builders for the test-suite: seeded weights (same generator as tools/make_golden.py), the oracle's
7-branch batch, golden loading and error metrics."""
from functools import lru_cache
from pathlib import Path

import torch

import convofusion_b200 as cf
from convofusion_b200.synthetic import randomize_, synthetic_clip
from oracle import convofusion_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
SCHED_KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")


def golden(name):
    return torch.load(GOLDEN / name, map_location="cpu", weights_only=True)


@lru_cache(maxsize=1)
def cpu_sampler():
    """fp32 sampler with the golden weights (seed 1234), parameters on the CPU (never run there)."""
    return randomize_(cf.ConvoFusionSampler(precision="fp32"), 1234)


@lru_cache(maxsize=1)
def state_dict():
    return {k: v.clone() for k, v in cpu_sampler().state_dict().items()}


def oracle_batch(syn):
    """Oracle 7*B batch from a synthetic clip (masks: True = pad)."""
    clip = dict(syn["clip"])
    clip["text_lsn_mask"] = ~clip["text_lsn_attn"].bool()
    clip["text_spk_mask"] = ~clip["text_spk_attn"].bool()
    return O.assemble_guidance_batch(state_dict(), clip, syn["uncond_text"], ~syn["uncond_text_attn"].bool())


def oracle_denoise(x, t, enc, masks):
    return O.denoiser_forward(state_dict(), x, t, enc, masks, prefix="denoiser.")


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    """max |a-b| / max |b|: the per-tensor relative error the north star's tolerances are stated in."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def frac_within(a, b, tol):
    """Fraction of elements with |a-b| <= tol * max|b| -- the north star's acceptance metric ("at least 90 % of
    fp32-reference outputs within tolerance at every DDIM step")."""
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() <= tol * b.abs().max()).double().mean())


def unbounded_windows(g_unbounded, n_streams):
    """Featurised per-window conditioning for golden("ref_loops.pt")["unbounded"]: the synthetic long batch cut by the
    package's own window bookkeeping (convofusion_b200.windows.slice_windows), with the stand-in T5 features of
    tools/pin_reference_loops.py.  Returns (list of clip dicts, uncond_text, uncond_text_attn)."""
    from convofusion_b200.synthetic import synthetic_long_batch, synthetic_text_features
    from convofusion_b200.windows import slice_windows
    long_batch = synthetic_long_batch(n_streams, g_unbounded["n_parts"], seed=g_unbounded["batch_seed"])
    syn = synthetic_clip(n_streams, seed=g_unbounded["uncond_clip_seed"], dyadic=True)
    uncond = (syn["uncond_text"], syn["uncond_text_attn"])

    def featurise(texts):
        f = [uncond if t == "-" * 10 else synthetic_text_features(t) for t in texts]
        return torch.stack([x[0] for x in f]), torch.stack([x[1] for x in f])

    wins = slice_windows(long_batch, featurise)
    for k, w in enumerate(wins):      # the texts the reference's process_text selected for the same windows
        assert w["texts"]["lsn"] == g_unbounded["texts_lsn"][k] and w["texts"]["spk"] == g_unbounded["texts_spk"][k]
    return wins, syn["uncond_text"], syn["uncond_text_attn"]
