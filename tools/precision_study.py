#!/usr/bin/env python
"""Where does the bf16 mode's trajectory error come from?  (task 1d.)

Runs the B = 1 DDIM-50 golden (tests/golden/sample_ddim50_clip.pt, outputs of the unmodified reference modules) and a
dyadic B = 8 DDIM-50 batch through the fp32 handle with its GEMM operands rounded in four ways
(cfb_set_fp32_tensor_cores):
    0  fp32 operands on the CUDA cores
    1  three-way bf16 split of both operands (fp32-accurate, the fp32 default)
    2  activations hi + lo (two bf16 terms, 16 mantissa bits) x weights rounded to bf16
    3  both operands rounded to bf16 -- the bf16 mode's GEMM rounding, everything else (LayerNorm, attention,
       softmax, residual, guidance, scheduler) in fp32
    4  like 2 for the GEMMs whose A operand is a LayerNorm output (qkv, both TimeBlock linears, scores / conditional
       queries, linear1, latent_proj), like 3 for the others (out_proj, values / fuser, linear2)
and through the real bf16 handle (bf16 GEMM operands AND bf16 attention memory / probabilities / intermediates), and
prints the relative L2 error of the latents after steps 1, 10, 25, 50 against the reference.

    python tools/precision_study.py > profiles/r02_precision_study.txt        (needs the B200)
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch

import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import SCHED_KW, golden, rel_err, state_dict

DEV = "cuda:0"
MODES = {0: "fp32 operands, CUDA cores", 1: "3 x 3 bf16 terms (fp32 default)", 2: "A hi+lo x W bf16",
         3: "A bf16 x W bf16 (GEMMs only)", 4: "A hi+lo only after LayerNorms"}


def sampler(precision):
    s = cf.ConvoFusionSampler(precision=precision)
    s.load_state_dict(state_dict())
    s = s.to(DEV).eval()
    s.scheduler = cf.DDIMScheduler(clip_sample=True, **SCHED_KW)
    s.num_inference_timesteps = 50
    return s


def run(s, syn, B, init, **kw):
    d = to_device(syn, DEV)
    enc, masks = s.encode_conditions(d["clip"], d["uncond_text"], d["uncond_text_attn"])
    _, rec, _ = s.sample(enc, masks, B, init.to(DEV), record=True, **kw)
    return rec.cpu()


def main():
    lib = _lib.lib()
    s32, s16 = sampler("fp32"), sampler("bf16")
    cases = []
    g = golden("sample_ddim50_clip.pt")
    cases.append(("B=1 monadic vs the reference modules' golden", synthetic_clip(1, seed=1235, dyadic=False), 1,
                  torch.randn(1, 16, 128, generator=torch.Generator().manual_seed(100)), g["record"]))
    syn8 = synthetic_clip(8, seed=41, dyadic=True)
    init8 = torch.randn(8, 16, 128, generator=torch.Generator().manual_seed(42))
    _lib.check(lib.cfb_set_fp32_tensor_cores(0))
    cases.append(("B=8 dyadic vs fp32 on the CUDA cores", syn8, 8, init8, run(s32, syn8, 8, init8)))
    steps = (0, 9, 24, 49)
    for title, syn, B, init, ref in cases:
        print(f"== {title}: relative L2 error of the latents after steps 1 / 10 / 25 / 50")
        for mode, name in MODES.items():
            _lib.check(lib.cfb_set_fp32_tensor_cores(mode))
            rec = run(s32, syn, B, init)
            print(f"   fp32 handle, GEMM operands {name:34s}: " + "  ".join(f"{rel_err(rec[i], ref[i]):.2e}" for i in steps))
        _lib.check(lib.cfb_set_fp32_tensor_cores(1))
        rec = run(s16, syn, B, init)
        print(f"   bf16 handle (6 branches)                                     : " + "  ".join(f"{rel_err(rec[i], ref[i]):.2e}" for i in steps))
        if "monadic" in title:
            rec = run(s16, syn, B, init, spk_is_uncond=True)
            print(f"   bf16 handle (speaker-only branch dropped: 5 branches)        : " + "  ".join(f"{rel_err(rec[i], ref[i]):.2e}" for i in steps))


if __name__ == "__main__":
    main()
