#!/usr/bin/env python
"""Row-block kernel (csrc/rowblock.cu) against the one-kernel-per-operator path, program kind by program kind.

Runs the bf16 sampling step with guidance scale 1 (so that bf16 rounding flips are not amplified 74x and a real
defect stands out) with cfb_set_rowblock(mask) for every mask and prints the deviation of the first-step and
last-step latents from mask 0; also the fp32 path as the yardstick.  `python tools/rb_check.py [B] [steps] [dyadic]`."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import state_dict


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def fault():
    lib = _lib.lib()
    out = (C.c_uint * 8)()
    lib.cfb_debug_rb_fault(out)
    return list(out)


def run(B=8, steps=2, dyadic=True, masks=(1, 2, 4, 7), scale=1.0, dev="cuda:0", verbose=True):
    res = {}
    samplers = {}
    for prec in ("bf16", "fp32"):
        s = cf.ConvoFusionSampler(precision=prec, num_inference_timesteps=steps)
        s.load_state_dict(state_dict())
        s = s.to(dev).eval()
        s.guidance_scale = scale
        samplers[prec] = s
    syn = to_device(synthetic_clip(B, seed=77 + B, dyadic=dyadic), dev)
    init = torch.randn(B, 16, 128, generator=torch.Generator().manual_seed(5 + B)).to(dev)
    sb = samplers["bf16"]
    enc, masks_ = sb.encode_conditions(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"])
    mono = not dyadic
    _, ref32, _ = samplers["fp32"].sample(enc, masks_, B, init, record=True, spk_is_uncond=mono)
    lib = _lib.lib()
    try:
        _lib.check(lib.cfb_set_rowblock(0))
        # the row-block programs write bf16 LayerNorm outputs: compare with the operator path in the same format
        _lib.check(lib.cfb_set_bf16_activation_f16(0))
        _, base, _ = sb.sample(enc, masks_, B, init, record=True, spk_is_uncond=mono)
        torch.cuda.synchronize()
        res["operator_vs_fp32"] = (rel(base[0], ref32[0]), rel(base[-1], ref32[-1]))
        for m in masks:
            _lib.check(lib.cfb_set_rowblock(m))
            for graph in (False, True):
                _, rec, _ = sb.sample(enc, masks_, B, init, record=True, spk_is_uncond=mono, use_graph=graph)
                torch.cuda.synchronize()
                res[(m, graph)] = (rel(rec[0], base[0]), rel(rec[-1], base[-1]), rel(rec[-1], ref32[-1]),
                                   bool(torch.isfinite(rec).all()))
    except Exception as exc:
        print("FAILED:", exc, "| row-block fault record {code, block, warp, stage, barrier, parity}:", fault())
        raise
    finally:
        lib.cfb_set_rowblock(0)
        lib.cfb_set_bf16_activation_f16(31)
    if verbose:
        print(f"B={B} steps={steps} dyadic={dyadic} guidance_scale={scale}")
        print(f"  operator path vs fp32: first {res['operator_vs_fp32'][0]:.3e} last {res['operator_vs_fp32'][1]:.3e}")
        for m in masks:
            for graph in (False, True):
                a, b, c, fin = res[(m, graph)]
                print(f"  rowblock mask {m} graph={int(graph)}: vs operator path first {a:.3e} last {b:.3e}; vs fp32 last {c:.3e}; finite {fin}")
    return res


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dy = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
    run(B, steps, dy)
